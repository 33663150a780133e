"""bench.py's flop accounting against the figures SURVEY.md 8(d) states (the numbers `roofline.achieved` is built from),
and the shape of the workload table (no GPU)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


@pytest.mark.parametrize("name,fwd_per_row", [
    ("C1", 1.08e4),        # "C1 1.08e4 fwd / 3.24e4 eval per row"
    ("C2", 3.43e5),        # "C2 3.43e5 / 1.03e6"
    ("T", 1.34e6),         # "T 1.34e6 / 4.03e6"
    ("C3", 1.77e7),        # "C3 1.77e7 / 5.31e7"
    ("C5", 2.11e7),        # "C5 2.11e7 fwd per row"
])
def test_forward_flops_per_row_match_the_survey(name, fwd_per_row):
    got = bench.workload_flops_per_row(bench.WORKLOADS[name])
    assert abs(got - fwd_per_row) / fwd_per_row < 5e-3


def test_row_layer_formula():
    # F(M, D_in, R) = 2 M D_in + M^2 + 2 M R + 2 M + R (M^2 + 2 M)
    assert bench.flops_per_row_layer(256, 8, 8) == 2 * 256 * 8 + 256 ** 2 + 2 * 256 * 8 + 2 * 256 + 8 * (256 ** 2 + 512)
    # the T evaluation: 3 F per row, 1e5 rows -> 4.03e11 flops -> 10.9 ms at 37 TFLOP/s
    per_eval = 3 * bench.workload_flops_per_row(bench.WORKLOADS["T"]) * 10_000 * 10
    assert abs(per_eval - 4.03e11) / 4.03e11 < 5e-3


def test_workload_table_matches_baseline_configs():
    import json
    base = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "BASELINE.json")))
    assert "north_star" in base
    w = bench.WORKLOADS
    assert (w["C1"]["L"], w["C1"]["M"], w["C1"]["D"], w["C1"]["N"], w["C1"]["S"]) == (2, 50, 1, 1000, 1)
    assert (w["C2"]["L"], w["C2"]["M"], w["C2"]["D"], w["C2"]["B"], w["C2"]["S"]) == (3, 128, 8, 10_000, 10)
    assert (w["T"]["L"], w["T"]["M"], w["T"]["D"], w["T"]["B"], w["T"]["S"]) == (3, 256, 8, 10_000, 10)
    assert (w["C3"]["L"], w["C3"]["M"], w["C3"]["D"], w["C3"]["B"], w["C3"]["S"], w["C3"]["kind"]) == \
        (5, 500, 16, 80_000, 16, "matern52")
    assert (w["C4"]["M"], w["C4"]["tasks"], w["C4"]["N"]) == (256, 8, 200_000)
    assert (w["C5"]["L"], w["C5"]["M"], w["C5"]["S"], bool(w["C5"].get("predict"))) == (3, 1024, 100, True)
