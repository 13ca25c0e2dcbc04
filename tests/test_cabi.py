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


def _header_prototypes():
    """{name: (return type text, [parameter type texts])} parsed from include/vmmt.h."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(vmmt_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = (ret, plist)
    return protos


def test_ctypes_signatures_match_header_prototypes():
    """Arity and pointer-ness / width class of every ctypes signature against the C prototype: a mismatch here is a
    silent stack corruption on the GPU box."""
    _lib_path()
    from variational_mmt_b200 import _lib
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)

    def kind(ctype_text):
        t = ctype_text
        if "*" in t:
            return "ptr"
        t = re.sub(r"\b(const|unsigned|[a-z_][a-z0-9_]*)$", lambda mm: mm.group(0), t)     # keep the text, drop nothing
        words = [w for w in re.split(r"\s+", t) if w]
        words = words[:-1] if len(words) > 1 and words[-1] not in ("int", "float", "long", "size_t", "int64_t", "uint64_t") \
            else words                                                                    # drop the parameter name
        base = " ".join(w for w in words if w != "const")
        return {"int": "i32", "float": "f32", "int64_t": "i64", "uint64_t": "i64", "size_t": "i64", "long long": "i64",
                "unsigned long long": "i64", "void": "void"}.get(base, base)

    cmap = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "i32", ctypes.c_int64: "i64",
            ctypes.c_uint64: "i64", ctypes.c_size_t: "i64", ctypes.c_float: "f32", ctypes.c_ulonglong: "i64", None: "void"}
    bad = []
    for name, (ret, plist) in protos.items():
        res, args = _lib.SIGNATURES[name]
        if len(args) != len(plist):
            bad.append(f"{name}: {len(plist)} parameters in the header, {len(args)} in _lib.SIGNATURES")
            continue
        for i, (a, ptxt) in enumerate(zip(args, plist)):
            want = kind(ptxt)
            got = "ptr" if (isinstance(a, type) and issubclass(a, ctypes._Pointer)) else cmap.get(a, "?")
            if want != got:
                bad.append(f"{name} arg {i} ('{ptxt}'): header {want}, ctypes {got}")
        want_r = kind(ret)
        got_r = cmap.get(res, "?")
        if want_r != got_r:
            bad.append(f"{name} return ('{ret}'): header {want_r}, ctypes {got_r}")
    assert not bad, "\n".join(bad)


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md (the reference-side binding guide) mentions every symbol include/vmmt.h declares."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "vmmt.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = []
    for sym in sorted(set(re.findall(r"\b(vmmt_[a-z0-9_]+)\s*\(", header))):
        if sym in doc:
            continue
        m = re.match(r"(vmmt_.*)_(fwd|bwd|wgrad)$", sym)
        if m and any((m.group(1) + suf) in doc for suf in ("_fwd/bwd", "_*")):
            continue
        missing.append(sym)
    assert not missing, missing
