"""simple_b200 -- B200-native (sm_100a) implementation of the SimPLe world-model hot path.

Public surface mirrors the reference (thomas-schillaci/SimPLe): NextFramePredictor, Adafactor, Trainer,
make_simulated_env.  All arithmetic runs in the hand-written CUDA kernels of csrc/ through the C ABI in
include/simple_b200.h; there is no CPU or PyTorch fallback.
"""
__version__ = '0.1'
