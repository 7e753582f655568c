"""Host accuracy of the FAST policy's branch-free primitives: decipher_b200/csrc/dcp_fastmath.cuh compiles with g++, so
the polynomial and the table-driven exp / log and the Newton reciprocal are checked against long-double libm without a GPU
(tests/cpp/fm_accuracy.cpp; the device versions are checked in tests/test_gpu_fastmath.py)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fast_exp_log_rcp_accuracy_on_the_host(tmp_path):
    exe = str(tmp_path / "fm_accuracy")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "cpp", "fm_accuracy.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    v = {out[i]: float(out[i + 1]) for i in range(0, len(out), 2)}
    print(v)
    assert v["exp_poly"] <= 1.5 and v["exp_tab"] <= 2.0
    assert v["log_poly"] <= 1.5 and v["log_tab"] <= 3.0
    assert v["log_tab_abs"] < 1e-18 and v["log_poly_abs"] < 1e-18      # |log x| <= 1e-3: absolute error
    assert v["rcp"] <= 1.5                                              # host stand-in for the hardware seed + two Newton steps
