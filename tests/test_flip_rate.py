"""Statistical pin of the K3 arithmetic spec against the REFERENCE sampler (parity gap 1).

The spec (oracle/spec_c: sdvar_spec_sample) fixes an exp polynomial and fixed-point sums where the reference leaves the
evaluation order to ATen (models/helpers.py:6-19), so a token may differ on a near-tie.  tools/flip_rate.py measured how
often over >= 1e6 rows (both filter settings x both logit scales) in the build container; the committed result is asserted
here, and a small live sample is re-measured whenever the reference is present."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESULT = os.path.join(ROOT, "profiles", "flip_rate_r02.json")
BOUND = 1e-5          # flips per sampled token row tolerated against the reference's torch sampler


def test_committed_flip_rate_is_below_bound():
    d = json.load(open(RESULT))
    assert d["total_rows"] >= 1_000_000
    assert {(c["top_k"], c["top_p"]) for c in d["configs"]} == {(0, 0.0), (900, 0.96)}
    assert {c["logit_scale"] for c in d["configs"]} == {0.05, 3.0}
    for c in d["configs"]:
        assert c["rows"] >= 250_000 and c["flips"] <= BOUND * c["rows"], c


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="the reference lives in the build container only")
def test_live_flip_sample_against_reference():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import flip_rate
    flips = rows = 0
    for ci in range(len(flip_rate.CONFIGS)):
        _, f, n = flip_rate.chunk((ci, 12345, 4))      # 4 images x 64 tokens per config
        flips += f
        rows += n
    assert rows == 4 * 256 and flips == 0
