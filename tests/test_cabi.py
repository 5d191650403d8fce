"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/srcnn_b200.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import _pkg

pkg = _pkg.load()


def header_symbols():
    src = open(pkg.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srcnn_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == pkg.exported_symbols()


def test_library_builds_and_exports_every_symbol():
    pkg.build()
    assert os.path.exists(pkg.LIB_PATH)
    L = C.CDLL(pkg.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SrcnnError, match="no CUDA device|CUDA"):
        pkg.Context(0)


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under the package may load or link it."""
    for root, _, files in os.walk(_pkg.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(root, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "libsrcnn_oracle", "libsrcnn_ref",
                               "srcnn_oracle.c", "oracle.loader"):
                    assert needle not in text, (root, f, needle)


def test_row_bands_cover_output_exactly():
    for out_h in (1, 7, 244, 4084, 1068):
        for n in (1, 2, 3, 4, 8):
            bands = pkg.row_bands(out_h, n)
            assert len(bands) == n
            rows = [r for (a, b) in bands for r in range(a, b)]
            assert rows == list(range(out_h))


def test_patch_shards_cover_exactly():
    for total in (0, 1, 5, 4096, 65536, 4097):
        for n in (1, 2, 4, 8):
            sh = pkg.patch_shards(total, n)
            assert sh[0][0] == 0 and sh[-1][1] == total
            assert all(sh[i][1] == sh[i + 1][0] for i in range(n - 1))
            sizes = [b - a for a, b in sh]
            assert max(sizes) - min(sizes) <= 1
