"""Pins the committed golden vectors to the reference's own code: re-runs oracle/gen_golden.py (which
imports /root/reference unmodified through oracle/ref_shim.py) in a subprocess and checks that the
fresh vectors equal the committed ones.  Only possible in the build container."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/adapteacher"), reason="reference tree not present")
def test_goldens_regenerate_from_reference(tmp_path, golden_dir):
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import oracle.gen_golden as g\n"
        "g.GOLDEN = %r\n"
        "sys.argv = ['x', 'ops', 'sampler']\n"
        "g.main()\n" % (ROOT, str(tmp_path)))
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for name in ("ops.npz", "sampler.npz"):
        a, b = np.load(tmp_path / name), np.load(os.path.join(golden_dir, name))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            if a[k].dtype.kind in "fc":
                np.testing.assert_allclose(a[k], b[k], atol=1e-6, rtol=1e-5, err_msg=k)
            else:
                assert np.array_equal(a[k], b[k]), k
