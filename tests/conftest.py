import ctypes
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


# ---- host emulation of the seq-attention kernel source (tests/emul): shared by the emulation and the module wiring tests ----
@pytest.fixture(scope="session")
def emulation(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("emul") / "libseq_attention_emul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-o", str(out),
                    os.path.join(ROOT, "tests", "emul", "seq_attention_emul.cpp")], check=True)
    lib = ctypes.CDLL(str(out))
    c_i32, c_ptr = ctypes.c_int32, ctypes.c_void_p
    lib.emul_seq_attention_fwd.restype, lib.emul_seq_attention_fwd.argtypes = c_i32, [c_ptr] * 9 + [c_i32] * 5 + [ctypes.c_float, ctypes.c_uint64] + [c_ptr] * 2
    lib.emul_seq_attention_bwd.restype, lib.emul_seq_attention_bwd.argtypes = c_i32, [c_ptr] * 12 + [c_i32] * 5 + [ctypes.c_float, ctypes.c_uint64] + [c_ptr] * 9
    return lib


@pytest.fixture()
def host_backend(emulation, monkeypatch):
    """Routes ops.SeqEdgeAttentionFn to the host emulation (same argument order as the C ABI, no stream)."""
    import torch

    def pointer(dtype):
        def get(t):
            if t is None:
                return None
            assert t.dtype == dtype and t.is_contiguous() and not t.is_cuda
            return t.data_ptr()
        return get

    def fwd(*args):
        assert emulation.emul_seq_attention_fwd(*args) == 0

    def bwd(*args):
        assert emulation.emul_seq_attention_bwd(*args) == 0

    from buglab_b200 import ops

    monkeypatch.setattr(ops, "_seq_attention_backend", lambda: (fwd, bwd, pointer(torch.float32), pointer(torch.int32)))
