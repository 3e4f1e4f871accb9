"""CPU test: the fp64 sum-product check-node evaluation of the CUDA path (csrc/kernels.cuh bp_check / tile4.cuh cn4_any: d exponentials,
fraction-valued forward / backward products on E = e^-|x|, d logarithms) emulated operation for operation (tests/bp_edomain_emul.h)
inside a copy of the C oracle, against the outputs of the UNMODIFIED reference decoder (tests/golden/*.npz, reference
src/decoding/decoder.cpp:11-78 with the pairwise box-plus of decoder.h:12-15).  The bar is north_star's: posteriors within 1e-4
relative, >= 99.99 % identical hard decisions; iteration counts are identical on every golden frame."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, H_FILE, H_IRREGULAR, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.fixture(scope="module")
def edomain_oracle(tmp_path_factory):
    import study_bp_edomain as S
    S.WORK = str(tmp_path_factory.mktemp("bp_edomain"))
    S.build()
    return os.path.join(S.WORK, "liboracle_e.so")


CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import oracle as O
O.LIB = %r
res = {}
for codefile, cases in ((%r, %r), (%r, %r)):
    code = O.Code(codefile)
    z = np.load(cases)
    names = sorted({k.split("/")[0] for k in z.files})
    for n in names:
        tags = [str(v) for v in z[n + "/names"]]   # (channel, decoder) in decode_cases.npz, (decoder, input kind) in irregular_cases.npz
        if "BP" not in tags or "BEC" in tags:
            continue
        it, et = int(z[n + "/cfg"][0]), bool(z[n + "/cfg"][1])
        out, co, its = code.decode(z[n + "/llr_in"], it, et, False)
        ref = z[n + "/llr_out"]
        rel = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-9)
        rel = np.where(np.isfinite(ref) & np.isfinite(out), rel, np.where(out == ref, 0.0, np.inf))
        res[n] = (float(rel.max()), float((co == z[n + "/co"]).mean()), bool(np.array_equal(its, z[n + "/iters"])), int(ref.shape[0]))
print(repr(res))
'''


def test_edomain_check_node_matches_reference_golden(edomain_oracle):
    src = CHILD % (ROOT, edomain_oracle, H_FILE, os.path.join(GOLDEN, "decode_cases.npz"), H_IRREGULAR, os.path.join(GOLDEN, "irregular_cases.npz"))
    r = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, env=dict(os.environ, EDOMAIN="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    res = eval(r.stdout.strip().splitlines()[-1])
    assert len(res) >= 6, res
    frames = 0
    for name, (rel, same, its_equal, n) in res.items():
        assert rel < 1e-4, (name, rel)
        assert same >= 0.9999, (name, same)
        assert its_equal, name
        frames += n
    assert frames >= 40
