"""the reference-compatible benchmark driver (driver/ref_table.cpp: the
reference's shape table, seed, generator, timing and print format, main.cu:24-80) calling the engine
through the reference's function-pointer type.  The reference's own driver never looks at a result;
this test does: every line's indices must be V0's on the same libc generator stream."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "nns-cuda_b200", "driver", "ref_table")
SHAPES = [(3, 1, 1024), (16, 1, 1024), (3, 1, 65536), (16, 1, 65536), (3, 1024, 1024),
          (16, 1024, 1024), (3, 1024, 65536), (16, 1024, 65536), (3, 1024, 1048576), (16, 1024, 1048576)]


def test_shape_table_is_the_references():
    src = open(os.path.join(ROOT, "nns-cuda_b200", "driver", "ref_table.cpp")).read()
    nums = [int(x) for x in re.search(r"samples\[\] = \{([^}]*)\}", src).group(1).replace("\n", " ").split(",")]
    assert [tuple(nums[i:i + 3]) for i in range(0, len(nums), 3)] == SHAPES  # main.cu:38-51


@pytest.mark.gpu
def test_ref_table_lines_and_results_equal_v0(oracle, tmp_path):
    assert os.path.exists(DRIVER), "build with make -C nns-cuda_b200"
    out = subprocess.run([DRIVER, "2", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("CudaCall 14,")]
    assert len(lines) == 2 * len(SHAPES)
    for ln, (k, m, n) in zip(lines, SHAPES):  # main.cu:76's format
        f = [x.strip() for x in ln[len("CudaCall "):].split(",")]
        assert (int(f[1]), int(f[2]), int(f[3])) == (k, m, n) and f[4].endswith("ms")
    # after the warm-up no line times context creation: the first line is a sub-millisecond search
    assert float(lines[0].split(",")[-1].strip()[:-2]) < 50.0, lines[0]
    total = exact = 0
    for i, (k, m, n, s, r) in enumerate(oracle.reference_table_inputs(SHAPES)):
        g = np.fromfile(os.path.join(str(tmp_path), f"results_{i}.bin"), dtype=np.int32)
        assert g.size == m
        v, _ = oracle.v0_omp(k, m, n, s, r)
        rep = oracle.check_tie_rule(k, m, n, s, r, g, v, 1e-5)
        assert rep["violations"] == 0 and rep["oracle_anomalies"] == 0, (k, m, n, rep)
        total += m
        exact += rep["exact_match_with_v0"]
    assert exact >= total - 2, (exact, total)
