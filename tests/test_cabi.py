"""CPU: the C-ABI library loads and exports every symbol include/vmmt.h declares; the product
package refuses to import without it (no fallback)."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vmmt.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vmmt_[a-z0-9_]+)\s*\(", text)))


def _lib_path():
    # load build.py by path: importing the package itself requires the library to exist already
    import importlib.util
    spec = importlib.util.spec_from_file_location("vmmt_build", os.path.join(ROOT, "variational_mmt_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(_lib_path())
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/vmmt.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    _lib_path()
    from variational_mmt_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    assert _lib.lib.vmmt_version() >= 100


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_fallback_when_library_missing(tmp_path):
    """Importing the package with the .so hidden must raise, not degrade to CPU / torch ops."""
    code = ("import sys, os\n"
            "so = os.path.join('variational_mmt_b200', 'libvmmt.so')\n"
            "os.rename(so, so + '.hidden')\n"
            "try:\n"
            "    try:\n"
            "        import variational_mmt_b200\n"
            "        print('IMPORTED')\n"
            "    except ImportError as e:\n"
            "        print('RAISED', 'no CPU' in str(e))\n"
            "finally:\n"
            "    os.rename(so + '.hidden', so)\n")
    _lib_path()
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert "RAISED True" in r.stdout, r.stdout + r.stderr


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "variational_mmt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
