"""tools/new_kernels_probe.py as a test: K1 lean with hot rows + user runs + wire input, group_by_user, the wire
packer / unpacker, the persistent epoch kernel on a 2-CTA cluster and the atomic epoch path, all against the oracle
on cases small enough that an out-of-bounds access would corrupt a checked result."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(300)
def test_new_kernels_probe():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "new_kernels_probe.py")], capture_output=True,
                         text=True, timeout=280)
    assert out.returncode == 0 and "probe ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
