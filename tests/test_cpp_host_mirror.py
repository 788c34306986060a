"""include/diffsol_b200.hpp, the C++ host-side mirror of the reference's OdeBuilder -> problem.bdf() -> solve_dense
surface over the C ABI: compiled with g++ against the in-tree library (tests/cpp/robertson_cpp.cpp).  Without a GPU
the program's host-side checks run and the solver construction fails loudly; on a GPU its output equals the Python
mirror's on the same inputs, bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def cpp_binary():
    from diffsol_b200 import build
    lib = build.build()
    out_dir = os.path.join(ROOT, "tests", "host_emu", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "robertson_cpp")
    src = os.path.join(ROOT, "tests", "cpp", "robertson_cpp.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("diffsol_b200.hpp", "diffsol_b200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in [src, lib] + hdrs):
        cxx = "/opt/gcc/bin/g++" if os.path.exists("/opt/gcc/bin/g++") else "g++"
        subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L", os.path.dirname(lib), "-ldiffsol_b200", "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return exe


def test_cpp_mirror_fails_loudly_without_a_device(cpp_binary):
    from diffsol_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([cpp_binary, "bdf", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "DiffsolError" in r.stderr      # host-side checks passed (codes 10-14), no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_cpp_mirror_equals_python_mirror(cpp_binary, method):
    import diffsol_b200 as dsb
    B = 6
    r = subprocess.run([cpp_binary, method, str(B)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [line.split() for line in r.stdout.strip().splitlines()]
    assert len(rows) == B
    p = np.array([[0.04 * (1.0 + 0.125 * b), 1.0e4, 3.0e7] for b in range(B)])
    solver = getattr(dsb.OdeBuilder().rhs_implicit("robertson_dae").p(p).rtol(1e-4).atol([1e-8, 1e-6, 1e-6]).build(), method)()
    ys = solver.solve_dense([0.4, 4.0, 40.0, 400.0, 4000.0, 40000.0])
    stats = solver.statistics_array()
    status = solver.status()
    for b, row in enumerate(rows):
        assert int(row[0]) == b and int(row[1]) == status[b] == 0
        assert [int(x) for x in row[2:15]] == stats[b, :13].tolist()
        assert [float(x) for x in row[15:18]] == ys[b, 5].tolist()
        assert float(row[18]) == 40000.0 and int(row[19]) == -1


@pytest.mark.gpu
def test_cpp_mirror_sensitivities_equal_python_mirror(cpp_binary):
    """OdeBuilder().sens_rtol().sens_atol() ... solve_dense_sensitivities through the C++ header, against the Python mirror."""
    import diffsol_b200 as dsb
    B = 5
    r = subprocess.run([cpp_binary, "sens", str(B)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [line.split() for line in r.stdout.strip().splitlines()]
    p = np.array([[0.1 * (1.0 + 0.25 * b), 1.0] for b in range(B)])
    solver = dsb.OdeBuilder().rhs_implicit("exp_decay").p(p).sens_rtol(1e-6).sens_atol([1e-6, 1e-6]).build().bdf_sens()
    ys, sens = solver.solve_dense_sensitivities([1.0, 2.0, 5.0])
    for b, row in enumerate(rows):
        assert int(row[0]) == b and int(row[1]) == 0
        assert [float(x) for x in row[2:5]] == [ys[b, 2, 0], sens[b, 2, 0, 0], sens[b, 2, 1, 0]]
        assert abs(float(row[3]) + 5.0 * np.exp(-5.0 * p[b, 0])) < 1e-5
