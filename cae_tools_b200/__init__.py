"""B200-native hot path for surftemp/cae_tools (ConvAEModel / UNET / VarAEModel
encoder-decoder forward+backward, MSE loss, Adam) behind the reference's own
Python API.  Host code is Python/PyTorch (memory, streams, torch.distributed);
all arithmetic on the hot path runs in hand-written sm_100a CUDA kernels reached
through the C-ABI library ``libcae_b200.so`` (see include/cae_b200.h)."""

VERSION = "0.1.0"
