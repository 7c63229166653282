"""CPU: the C-ABI library loads, exports every symbol include/nerf_b200.h declares, and fails loudly
without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

import nerf_b200
from nerf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(names=("nerf_b200.h", "nerf_b200_debug.h")):
    out = set()
    for name in names:
        src = open(os.path.join(ROOT, "include", name)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        out |= set(re.findall(r"\b(nb2_[a-z0-9_]+)\s*\(", src))
    return sorted(out)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `make` (or __graft_entry__.build()) first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nerf_b200.h but not exported"


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.nb2_version() == 200


def test_render_params_struct_matches_header():
    # 2 int, 4 float, 2 int, u64, i64, 4 int, 8 pointers -> 128 bytes with natural alignment
    assert ctypes.sizeof(_lib.RenderParams) == 128
    p = _lib.RenderParams(n_coarse=64, n_fine=128, precision=_lib.PREC_BF16X3)
    assert _lib.load().nb2_render_workspace_bytes(1000, ctypes.byref(p)) == 2 * 256000 + 512000 + 256
    p.precision = _lib.PREC_FP32
    assert _lib.load().nb2_render_workspace_bytes(1000, ctypes.byref(p)) == 2 * 256000 + 512000 + 256 + 2048000


def test_ctypes_structs_match_the_header_as_compiled_by_gcc(tmp_path):
    """Every struct that crosses the C ABI by pointer: sizeof and the offsets of the last fields as a C compiler lays out
    include/nerf_b200.h must equal the ctypes mirrors in nerf_b200/_lib.py (a field added on one side only shifts everything
    after it silently)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "nerf_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(nb2_render_params), sizeof(nb2_gemm_desc), offsetof(nb2_gemm_desc, split_stride),
         offsetof(nb2_gemm_desc, a_rowsum_out), offsetof(nb2_gemm_desc, a_rowsum_stride), sizeof(nb2_reduce_desc), offsetof(nb2_reduce_desc, out),
         sizeof(nb2_to_bf16_desc), offsetof(nb2_to_bf16_desc, lo));
  return 0;
}
""")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    G, R, T = _lib.GemmDesc, _lib.ReduceDesc, _lib.ToBf16Desc
    want = [ctypes.sizeof(_lib.RenderParams), ctypes.sizeof(G), G.split_stride.offset, G.a_rowsum_out.offset, G.a_rowsum_stride.offset,
            ctypes.sizeof(R), R.out.offset, ctypes.sizeof(T), T.lo.offset]
    assert got == want, (got, want)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(nerf_b200.NB2Error):
        _lib.handle()
    out = ctypes.c_void_p()
    rc = _lib.load().nb2_create(ctypes.byref(out), 0)
    assert rc < 0 and b"no CPU path" in _lib.load().nb2_last_error()
    net = nerf_b200.MipNeRF(10, 4)
    with pytest.raises(nerf_b200.NB2Error):
        net.forward(torch.zeros(2, 4, 6))
    with pytest.raises(nerf_b200.NB2Error):
        nerf_b200.positional_encoding(torch.zeros(4, 3), 10)


def test_state_dict_keys_match_reference():
    """Checkpoint compatibility surface."""
    m = nerf_b200.MipNeRF(10, 4, 256)
    p = nerf_b200.ProposalNetwork(10, 256)
    from oracle.nerf_oracle import NERF_KEYS, PROPOSAL_KEYS, layer_shapes
    assert [k[:-7] for k in m.state_dict() if k.endswith(".weight")] == NERF_KEYS
    assert [k[:-7] for k in p.state_dict() if k.endswith(".weight")] == PROPOSAL_KEYS
    assert [tuple(v.shape) for k, v in m.state_dict().items() if k.endswith(".weight")] == layer_shapes("nerf")
    assert [tuple(v.shape) for k, v in p.state_dict().items() if k.endswith(".weight")] == layer_shapes("proposal")
    assert sum(x.numel() for x in m.parameters()) == 530052 and sum(x.numel() for x in p.parameters()) == 214017
