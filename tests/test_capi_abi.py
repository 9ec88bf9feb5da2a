"""CPU test: the C-ABI shared library loads without a GPU and exports every symbol include/hybvio_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hybvio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hybvio_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "build the library first: make (or __graft_entry__.build())"
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/hybvio_b200.h but not exported: {missing}"


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the context must fail loudly (HV_ERR_NO_DEVICE), never fall back to a CPU path."""
    import torch
    from hybvio_b200 import capi
    lib = capi.load()
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = lib.hv_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.hv_last_error()


def test_version_string():
    from hybvio_b200 import capi
    assert b"sm_100a" in capi.load().hv_version()


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/hybvio_b200.h must compile as C (no C++-isms, no torch / CUDA types)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "hybvio_b200.h"\nint main(void) { hv_camera_model c; hv_visual_update_params p; hv_track_result r; hv_ekf_op o; '
                   '(void)c; (void)p; (void)r; (void)o; return 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ctypes_structs_match_the_header(tmp_path):
    """hybvio_b200/capi.py mirrors the structs of include/hybvio_b200.h by hand: sizes and field offsets must agree with what
    the C compiler lays out."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from hybvio_b200 import capi
    pairs = {"hv_camera_model": capi.CameraModel, "hv_track_obs": capi.TrackObs, "hv_track_model": capi.TrackModel,
             "hv_visual_update_params": capi.VisualUpdateParams, "hv_track_result": capi.TrackResult, "hv_ekf_op": capi.EkfOp,
             "hv_ekf_params": capi.EkfParams, "hv_lk_job": capi.LkJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hybvio_b200.h"', 'int main(void) {']
    for cname, py in pairs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in py._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for ln in out.splitlines():
        a, b, c = ln.split()
        got[(a, b)] = int(c)
    for cname, py in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(py), cname
        for fname, _ in py._fields_:
            assert got[(cname, fname)] == getattr(py, fname).offset, (cname, fname)
