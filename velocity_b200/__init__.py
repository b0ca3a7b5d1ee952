"""velocity_b200 -- B200-native (sm_100a) SFM speed-estimation hot path behind the reference's
Python function API (utils.KLT / utils.NLS / utils.MSV / utils.transforms of ultralytics/velocity).

The compute lives in velocity_b200/csrc (hand-written CUDA, C-ABI in include/velocity_b200.h);
the modules here are the thin host-side mirror of the reference's function signatures.  There is
no CPU fallback: every accelerated entry point raises if libvelocity_b200.so is missing.
"""
__version__ = "0.1.0"
