"""buglab_b200 — B200-native kernels (C ABI in ``include/buglab_b200.h``) and their PyTorch bindings for the
gnn-mlp message-passing hot path of BugLab.  No CPU fallback: everything here needs the in-tree
``libbuglab_b200.so`` and CUDA tensors."""
import torch as _torch

# fp32 parity (<=1e-4 vs the reference CPU path) forbids silent TF32 in the library GEMMs
_torch.backends.cuda.matmul.allow_tf32 = False
_torch.backends.cudnn.allow_tf32 = False

from . import _lib  # noqa: E402,F401
