"""The C-ABI library loads and exports every symbol include/taco_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "taco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(taco_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tacotron_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in taco_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_string():
    from tacotron_b200 import _lib
    lib = _lib.lib()
    assert lib.taco_version() == 100
    assert isinstance(lib.taco_last_error(), bytes)


def test_struct_sizes_match_header_layout():
    """ctypes mirrors must have the C layout: compile a tiny probe with gcc against the header."""
    import subprocess, tempfile
    from tacotron_b200 import _lib
    probe = r'''
#include <stdio.h>
#include "taco_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(taco_linear_desc), sizeof(taco_decoder_weights), sizeof(taco_decoder_args),
                       sizeof(taco_gemm_desc), sizeof(taco_decoder_bwd_args)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c"); exe = os.path.join(d, "p")
        open(c, "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b, cc, g, db = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_lib.LinearDesc) == a
    assert ctypes.sizeof(_lib.DecoderWeights) == b
    assert ctypes.sizeof(_lib.DecoderArgs) == cc
    assert ctypes.sizeof(_lib.GemmDesc) == g
    assert ctypes.sizeof(_lib.DecoderBwdArgs) == db


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from tacotron_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_lib.TacoError):
        _lib.lib()


def test_decoder_sizes_host_only():
    from tacotron_b200 import _lib
    lib = _lib.lib()
    # packed floats (decoder v4) = 32 column slices x the per-slice MMA A-fragment blocks [warps][k-tile slots][32 lanes][FL]
    # (FL = 2: 8 weight columns, FL = 4: 16), rounded to 64, + the fused weight-only products of the packed tail
    def expect(r):
        out = 80 * r
        blk = lambda nw, kpw, fl: nw * kpw * 32 * fl
        per_slice = (blk(16, 2, 2) + blk(16, 3, 2)                 # IN: s(t-1) half, [ctx | p2] half
                     + 3 * (2 * blk(16, 2, 4) + 2 * blk(16, 2, 2)) # 3 x (gates h/x halves, candidate x / r*h halves)
                     + 2 * blk(16, 2, 4)                           # y tile, [q | p1'] tile
                     + blk(16, 2, 2) + blk(8, 4, 2))               # teacher pre-net layer 1, pre-net layer 2
        sl = (32 * per_slice + 63) // 64 * 64
        tail = (256 + 128) * 256 + out * 256 + 256 * 256 + 256 * 512 + 256 + 256 + 256
        return sl + tail
    assert lib.taco_decoder_packed_bytes(5) // 4 == expect(5)
    assert lib.taco_decoder_packed_bytes(2) // 4 == expect(2)
    assert lib.taco_decoder_workspace_bytes(32, 128, 200, 5) > 0
