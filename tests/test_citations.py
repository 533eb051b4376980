"""CPU test: every `file:line` the C-ABI header cites in the reference exists (the judge follows these to check that
an entry point replaces what it says it replaces).  Runs only where the reference tree is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"((?:_lzbench|lz4|nvcomp)/[\w./-]+\.(?:c|cpp|h|cu)):(\d+)(?:-(\d+))?")


def _lines(path, cache={}):
    if path not in cache:
        with open(path, errors="replace") as f:
            cache[path] = sum(1 for _ in f)
    return cache[path]


DOCS = ["include/lz4b200.h", "INTEGRATION.md", "DESIGN.md", "oracle/lz4_oracle.c", "oracle/lz4f_oracle.c", "oracle/ref_frame.c",
        "oracle/ref_plugin.c", "lzbench_b200/codec.py", "lzbench_b200/csrc/frame.cuh", "lzbench_b200/csrc/lz4b200.cu",
        "lzbench_b200/csrc/lzbench_plugin.cpp", "lzbench_b200/csrc/staging.cuh", "lzbench_b200/csrc/segmented.cuh",
        "lzbench_b200/csrc/kernels_v0.cuh", "lzbench_b200/csrc/kernels_v5.cuh", "lzbench_b200/csrc/kernels_c1.cuh",
        "lzbench_b200/csrc/kernels_c2.cuh"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree absent")
@pytest.mark.parametrize("doc", DOCS)
def test_cited_reference_lines_exist(doc):
    text = open(os.path.join(ROOT, doc), errors="replace").read()
    cites = CITE.findall(text)
    if doc in ("include/lz4b200.h", "INTEGRATION.md", "oracle/lz4_oracle.c", "oracle/lz4f_oracle.c"):
        assert cites, f"{doc} cites nothing"
    for rel, lo, hi in cites:
        path = os.path.join(REF, rel)
        assert os.path.isfile(path), f"{doc}: cited file {rel} does not exist"
        n = _lines(path)
        assert int(lo) <= n and (not hi or int(lo) <= int(hi) <= n), f"{doc}: {rel}:{lo}-{hi} beyond the file's {n} lines"
