"""CPU tests of the drop-in boundary: libompimg_b200.so loads without a GPU, exports exactly the
symbols include/ompimg_b200.h declares (including the two frozen reference prototypes), the Python
bindings cover them all, and without a device every call fails loudly instead of falling back."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "ompimg_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"^OMPIMG_API[^;(]*?\b(\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_frozen_prototypes():
    text = HEADER.read_text()
    # monolithic/include/parallel_to_grayscale.h:4 and monolithic/include/sobel.h:3-5, verbatim argument lists
    assert re.search(r"void convert_to_grayscale\(unsigned char \*data, int width, int height, int channels\);", text)
    assert re.search(r"void sobel_edge\(const unsigned char \*src, unsigned char \*dst, int width, int height\);", text)
    for hdr, sym in (("parallel_to_grayscale.h", "convert_to_grayscale"), ("sobel.h", "sobel_edge")):
        assert sym in (ROOT / "include" / hdr).read_text()


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    syms = declared_symbols()
    assert len(syms) >= 60
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ompimg_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(syms), (exported ^ set(syms))
    assert set(pkg._SIGS) == set(syms)


def test_no_torch_and_no_oracle_in_the_product():
    """The product path must not route through the oracle or any CPU fallback."""
    for p in list((ROOT / "openmp-parallel-computing_b200").rglob("*.py")) + list((ROOT / "openmp-parallel-computing_b200" / "csrc").glob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h"):
            t = p.read_text()
            assert "import oracle" not in t and "liboracle" not in t and "oracle/" not in t, p
            assert "import torch" not in t, p


def test_fails_loudly_without_a_gpu(pkg):
    L = pkg.lib()
    if L.ompimg_device_count() > 0:
        pytest.skip("a GPU is visible here")
    img = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(pkg.OmpimgError) as e:
        pkg.convert_to_grayscale(img)
    assert "no usable CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    with pytest.raises(pkg.OmpimgError):
        pkg.sobel_edge(np.zeros((8, 8), np.uint8))
    with pytest.raises(pkg.OmpimgError):
        pkg.init(0)


def test_frozen_entry_point_aborts_without_a_gpu(pkg):
    """convert_to_grayscale returns void: the only error channel is a non-zero exit."""
    if pkg.lib().ompimg_device_count() > 0:
        pytest.skip("a GPU is visible here")
    code = ("import ctypes, numpy as np; L = ctypes.CDLL(%r); a = np.zeros(48, np.uint8);"
            "L.convert_to_grayscale(ctypes.c_void_p(a.ctypes.data), 4, 4, 3)") % str(pkg.LIB_PATH)
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "convert_to_grayscale failed" in r.stderr


def test_reference_drivers_link_against_the_library(tmp_path, pkg):
    """A C caller written like the reference's main.c links with nothing but -lompimg_b200."""
    src = tmp_path / "caller.c"
    src.write_text('#include "parallel_to_grayscale.h"\n#include "sobel.h"\n#include "ompimg_b200.h"\n'
                   "int main(int argc, char **argv) { unsigned char px[48] = {0}, e[16] = {0};\n"
                   " if (argc > 1) { convert_to_grayscale(px, 4, 4, 3); sobel_edge(px, e, 4, 4); }\n"
                   " return ompimg_device_count() < 0; }\n")
    exe = tmp_path / "caller"
    r = subprocess.run(["gcc", "-O2", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{pkg.LIB_PATH.parent}",
                        "-lompimg_b200", f"-Wl,-rpath,{pkg.LIB_PATH.parent}", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_integration_makefiles_build_the_reference_drivers(tmp_path, pkg):
    """integration/Makefile and integration/Makefile_with_sobel (the corrected replacement for the reference's broken
    monolithic/Makefile_with_sobel:11) build the reference's UNMODIFIED drivers against the library."""
    ref = Path("/root/reference/monolithic")
    if not (ref / "src" / "main_with_sobel.c").exists():
        pytest.skip("the reference tree is not on this box")
    for mk, exe in (("Makefile", "grayscale"), ("Makefile_with_sobel", "grayscale_sobel")):
        r = subprocess.run(["make", "-C", str(ref), "-f", str(ROOT / "integration" / mk), f"BIN_DIR={tmp_path}", f"B200={pkg.LIB_PATH.parent}"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        syms = subprocess.run(["nm", "-D", "--undefined-only", str(tmp_path / exe)], capture_output=True, text=True, check=True).stdout
        assert "convert_to_grayscale" in syms and ("sobel_edge" in syms) == (exe == "grayscale_sobel")
