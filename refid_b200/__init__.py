"""refid_b200: B200-native (sm_100a) backend for REFID's FinalBidirectionAttenfusion forward/backward."""
__version__ = "0.1.0"
