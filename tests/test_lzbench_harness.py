"""Drop-in check through the reference's OWN harness (BASELINE configs[0]: `lzbench -elz4 -b64` on a 64 MiB
synthetic mixed file): oracle/_ref/lzbench_b200 is the reference lzbench.cpp/compressors.cpp compiled from
the reference tree with integration/patch_lzbench.py applied (the comp_desc[] row + batch hooks) and linked
against liblz4b200.so; oracle/_ref/lzbench_ref is the unmodified harness.  Built by integration/build_lzbench.sh
in the build container; the binaries travel to the GPU box."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import lzbench_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lzbench_ref")
B200_BIN = os.path.join(ROOT, "oracle", "_ref", "lzbench_b200")

pytestmark = pytest.mark.gpu


def run(binary, args):
    r = subprocess.run([binary] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout + r.stderr


def rows(text):
    """codec name -> (compress MB/s, decompress MB/s or 'ERROR', compressed size)"""
    out = {}
    for line in text.splitlines():
        m = re.match(r"^(\S+(?: \S+)?)\s+(\d+) MB/s\s+(\d+ MB/s|ERROR)\s+(\d+)\s+([\d.]+)", line)
        if m:
            out[m.group(1)] = (int(m.group(2)), m.group(3), int(m.group(4)))
    return out


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(B200_BIN)), reason="oracle/_ref harness binaries not built")
@pytest.mark.parametrize("chunk_kb", [64, 4])
def test_lz4_b200_row_next_to_lz4(chunk_kb):
    n = 64 << 20
    data = L.corpus("mixed", n, seed=1)
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        data.tofile(f)
        path = f.name
    try:
        ref = rows(run(REF_BIN, ["-elz4", f"-b{chunk_kb}", "-t1,1", "-x", path]))
        ours = rows(run(B200_BIN, ["-elz4/lz4_b200", f"-b{chunk_kb}", "-t1,1", "-x", path]))
    finally:
        os.unlink(path)
    assert "lz4 1.9.3" in ref and "lz4 1.9.3" in ours and "lz4_b200 0.1" in ours, (ref, ours)
    # the patched harness leaves the stock row untouched
    assert ours["lz4 1.9.3"][2] == ref["lz4 1.9.3"][2]
    # round trip verified by the harness's own memcmp (lzbench.cpp:433-458 prints ERROR otherwise)
    assert ours["lz4_b200 0.1"][1] != "ERROR"
    assert ours["lz4_b200 0.1"][2] <= 1.03 * ref["lz4 1.9.3"][2]
    print("lzbench rows:", ours)
