"""mtts-b200: B200-native (sm_100a) implementation of the MOSS-TTSD batched generation hot path.

Drop-in mirrors of the reference's Python entry points live in
  moss_ttsd_b200.modeling_asteroid   (AsteroidTTSConfig / AsteroidTTSInstruct.generate)
  moss_ttsd_b200.generation_utils    (process_batch and the prompt-grid helpers)
  moss_ttsd_b200.xy_tokenizer        (XY_Tokenizer.encode / decode, ResidualVQ)
All arithmetic on the path runs in hand-written CUDA behind the C-ABI of include/mtts.h (libmtts.so);
there is no CPU or PyTorch fallback: importing `moss_ttsd_b200._lib` without the library raises.
"""
__version__ = "0.1.0"
