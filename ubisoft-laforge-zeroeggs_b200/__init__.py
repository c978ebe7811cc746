"""zeggs_b200 -- B200-native (sm_100a) implementation of the ZeroEGGS audio->gesture hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all compute on
the path runs in hand-written CUDA behind the C ABI declared in include/zeggs_b200.h
(libzeggs_b200.so, built in-tree by __graft_entry__.build()).  There is no CPU fallback:
importing the compute modules without the built library raises.
"""
__version__ = "0.1.0"
