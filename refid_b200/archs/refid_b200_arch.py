"""The plug-in file.  Copied (or symlinked) into the reference tree as `basicsr/models/archs/refid_b200_arch.py`, it is
picked up by the reference's own filename scan (basicsr/models/archs/__init__.py:9-18) and `define_network`
(:43-46) then returns THIS class for `network_g.type: FinalBidirectionAttenfusion` (see INTEGRATION.md for the one-line
removal of the stock definition so the scan finds exactly one).  Nothing else in the reference changes."""
from refid_b200.arch import FinalBidirectionAttenfusion  # noqa: F401

__all__ = ["FinalBidirectionAttenfusion"]
