"""CPU: the C-ABI library loads, exports every symbol include/nbody_b200.h declares, its structs
have the layout the Python mirror assumes, the file formats round-trip, and compute entry points
fail loudly (NB_ERR_NO_GPU) when no GPU is present — there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from conftest import ROOT, case_path, golden_lines

HEADER = os.path.join(ROOT, "include", "nbody_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(nb):
    L = nb.lib()
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), "libnbody_b200.so does not export %s" % name
    assert sorted(nb.ABI_SYMBOLS) == decl


def test_exports_are_c_linkage(nb):
    out = subprocess.check_output(["nm", "-D", "--defined-only", nb.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    for name in declared_symbols():
        assert name in exported


def test_struct_layout_matches_header(nb, tmp_path):
    """Compile a tiny C program against the header and compare sizeof / offsetof with ctypes."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nbody_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(nb_system), sizeof(nb_events), sizeof(nb_answer),'
                   'offsetof(nb_events, reach_step), offsetof(nb_answer, q3_cost), offsetof(nb_answer, pair_interactions));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(nb.NbSystem), C.sizeof(nb.NbEvents), C.sizeof(nb.NbAnswer), nb.NbEvents.reach_step.offset,
            nb.NbAnswer.q3_cost.offset, nb.NbAnswer.pair_interactions.offset]
    assert got == want


def test_version_and_strerror(nb):
    L = nb.lib()
    assert b"sm_100a" in L.nb_version()
    assert b"no CPU fallback" in L.nb_strerror(nb.NB_ERR_NO_GPU)
    assert L.nb_kernel_launches() >= 0


def test_read_input_matches_python_float_parsing(nb):
    """nbody.cc:22-39: every number is parsed correctly rounded (strtod == Python float())."""
    for case in ("b20", "b200", "b1024"):
        s = nb.read_input(case_path(case))
        lines = open(case_path(case)).read().split("\n")
        n, planet, asteroid = [int(x) for x in lines[0].split()]
        assert (s.n, s.planet, s.asteroid) == (n, planet, asteroid)
        for i in (0, 1, n // 2, n - 1):
            f = lines[1 + i].split()
            vals = [float(x) for x in f[:7]]
            got = [s.q[i], s.q[i + n], s.q[i + 2 * n], s.v[i], s.v[i + n], s.v[i + 2 * n], s.m[i]]
            assert got == vals
            assert bool(s.is_device[i]) == (f[7] == "device")
    assert nb.read_input(case_path("b80")).devices == [76, 77, 78, 79]


def test_write_output_is_the_reference_format(nb, tmp_path):
    """nbody.cc:41-49: std::scientific, 17 significant digits == the goldens' bytes."""
    for case in ("b20", "b200", "b1024"):
        g = golden_lines(case)
        p = tmp_path / (case + ".out")
        nb.write_output(str(p), g["min_dist"], g["hit_time_step"], g["gravity_device_id"], g["missile_cost"])
        assert p.read_text() == g["text"]
        assert nb.format_output(g["min_dist"], g["hit_time_step"], g["gravity_device_id"], g["missile_cost"]) == g["text"]


def test_io_errors_are_codes_not_exceptions_across_the_abi(nb, tmp_path):
    n, p, a = C.c_int(), C.c_int(), C.c_int()
    assert nb.lib().nb_read_header(b"/nonexistent/file.in", C.byref(n), C.byref(p), C.byref(a)) == nb.NB_ERR_IO
    bad = tmp_path / "bad.in"
    bad.write_text("3 0 1\n1 2 3 4 5 6 7 planet\n1 2 3\n")
    with pytest.raises(nb.NbodyError) as e:
        nb.read_input(str(bad))
    assert e.value.code == nb.NB_ERR_IO
    assert nb.lib().nb_write_output(b"/nonexistent/dir/x.out", 1.0, 1, 1, 1.0) == nb.NB_ERR_IO


def test_read_input_rejects_out_of_range_planet_or_asteroid(nb, tmp_path):
    """nbody.cc:118-120 indexes its arrays with the header's planet / asteroid unchecked; here a bad header is an I/O error."""
    body = "0 0 0 0 0 0 1e20 star\n" * 3
    for hdr in ("3 3 1", "3 0 -1", "3 7 7"):
        p = tmp_path / "bad.in"
        p.write_text(hdr + "\n" + body)
        with pytest.raises(nb.NbodyError) as e:
            nb.read_input(str(p))
        assert e.value.code == nb.NB_ERR_IO
    p = tmp_path / "ok.in"
    p.write_text("3 2 1\n" + body)
    assert nb.read_input(str(p)).planet == 2


def test_argument_validation(nb):
    L = nb.lib()
    q = np.zeros(3)
    assert L.nb_run_steps(0, 0, 0, None, None, None, None, 0, 1) == nb.NB_ERR_ARG
    assert L.nb_run_steps(0, 7, 1, nb._d(q), nb._d(q), nb._d(q), nb._u(np.zeros(1, np.uint8)), 0, 1) == nb.NB_ERR_ARG
    assert L.nb_large_scratch_bytes(65536, 8192) == 128 * 3 * 8192 * 8  # 16 i-blocks x 128 j-splits fill 148 SMs
    assert L.nb_large_scratch_bytes(65536, 65536) == 32 * 3 * 65536 * 8


def test_compute_without_gpu_fails_loudly(nb):
    if nb.device_count() > 0:
        pytest.skip("a GPU is present")
    s = nb.read_input(case_path("b20"))
    with pytest.raises(nb.NbodyError) as e:
        nb.solve(s)
    assert e.value.code == nb.NB_ERR_NO_GPU
    with pytest.raises(nb.NbodyError):
        nb.run_steps(0, 1, s.n, s.q.copy(), s.v.copy(), s.m, s.is_device)
    with pytest.raises(nb.NbodyError):
        nb.Trajectory(s, nb.KIND_Q1)


def test_hw5_cli_requires_two_arguments(nb):
    """hw5.cu:533-535 / nbody.cc:92-94: argc != 3 throws std::runtime_error -> abnormal exit."""
    r = subprocess.run([nb.HW5_PATH], capture_output=True)
    assert r.returncode != 0 and b"must supply 2 arguments" in r.stderr


def test_synthetic_system_is_reproducible(nb):
    a, b = nb.synthetic_system(4096, seed=42), nb.synthetic_system(4096, seed=42)
    assert np.array_equal(a.q, b.q) and np.array_equal(a.m, b.m)
    assert a.devices == [4092, 4093, 4094, 4095] and (a.planet, a.asteroid) == (0, 1)
    assert 1e20 <= a.m.min() and a.m.max() <= 1e30
