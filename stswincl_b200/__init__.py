"""stswincl_b200 -- B200-native (sm_100a) drop-ins for the two STswinCL hot paths:
the spatio-temporal shifted-window attention block and the pixel contrastive loss."""
__version__ = "0.1.0"
