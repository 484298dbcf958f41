"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/mt_b200.h declares; argument validation works without a GPU (no kernel is launched here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from metatransformer_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    text = (ROOT / "include" / "mt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mt_b200.h but not exported by libmt_b200.so"
    from metatransformer_b200._lib import EXPORTED_SYMBOLS
    assert sorted(EXPORTED_SYMBOLS) == declared


def test_version_and_error_string(lib):
    assert lib.mt_version() >= 100
    assert isinstance(lib.mt_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # all of these must be rejected on the host before any CUDA call
    assert lib.mt_gemm_bf16(None, None, None, None, None, 128, 256, 64, 0, 0, None) == -1
    assert b"null" in lib.mt_last_error()
    buf = ctypes.create_string_buffer(4096)
    p = (ctypes.addressof(buf) + 15) & ~15
    assert lib.mt_gemm_bf16(p, p, None, None, p, 0, 256, 64, 0, 0, None) == -1          # M = 0
    assert lib.mt_gemm_bf16(p, p, None, None, p, 128, 256, 60, 0, 0, None) == -1         # K % 8
    assert lib.mt_gemm_bf16(p + 2, p, None, None, p, 128, 256, 64, 0, 0, None) == -3     # alignment
    assert lib.mt_gemm_bf16(p, p, None, None, p, 128, 256, 64, 7, 0, None) == -4         # unknown activation
    assert lib.mt_layernorm(p, 0, p, p, p, 4, 12, 1e-5, None) == -1                       # d % 8
    assert lib.mt_layernorm(p, 5, p, p, p, 4, 16, 1e-5, None) == -2                       # dtype
    assert lib.mt_attention_fwd(p, p, 1, 16, 2, 72, 0.2, None) == -4                      # head_dim 72 unsupported
    assert b"head_dim" in lib.mt_last_error()
    assert lib.mt_im2row(p, 1, p, 1, 3, 1, 8, 8, 1, 16, 16, 1, 16, 16, None) == -1        # kernel > input
    assert lib.mt_cast(p, 1, p, 9, 16, None) == -2
    assert lib.mt_encoder_create(None, 768, 12, 3072, 12, 1e-5, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from metatransformer_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("MT_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MtError, match="no CPU / eager fallback"):
        _lib.load()


def test_header_is_plain_c_and_product_never_imports_the_oracle(tmp_path):
    """The boundary is a C ABI: the header must compile as C (no C++ / torch types); and the product package must not
    reference oracle/ anywhere (the oracle is test infrastructure)."""
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    gcc = shutil.which("gcc")
    if gcc:
        src = tmp_path / "use_header.c"
        src.write_text('#include "mt_b200.h"\n'
                       "int probe(void) { mt_gemm_opts o; o.act = 0; mt_block_weights w; (void)w; (void)o;\n"
                       "  return mt_version() + (int)MT_ERR_CUDA + (int)MT_DT_F16 + (int)MT_POOL_FIRST_TWO; }\n")
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-I{root / 'include'}", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    offenders = [str(p.relative_to(root)) for p in (root / "metatransformer_b200").rglob("*")
                 if p.suffix in (".py", ".cu", ".cuh", ".h") and "lib" not in p.parts
                 and ("from oracle" in p.read_text(errors="ignore") or "import oracle" in p.read_text(errors="ignore"))]
    assert offenders == [], offenders
