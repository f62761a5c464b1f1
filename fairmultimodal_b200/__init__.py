"""fairmultimodal_b200 -- B200-native (sm_100a) kernels behind the FAME hot path (10_FAME.py of the reference)."""
__version__ = "0.1.0"
