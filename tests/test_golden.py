"""Golden replay: every case of tests/golden/manifest.json (inputs + outputs of the UNMODIFIED reference binary, made by
tests/golden/make_golden.py) is run through

  * the CPU restatement (oracle/_ref/lr2rmats_port: oracle port + the product's host readers/emitters)  -- CPU, always;
  * the product CLI (lr2rmats_b200/host/lr2rmats-b200: CUDA library through the C ABI)                 -- `-m gpu`.

Outputs must be byte-identical (BAM outputs are compared after decompression: the deflate bytes depend on zlib).
"""
import gzip
import json
import os
import subprocess

import pytest

from tests import oracle_port as op

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
CASES = [(c, k) for c in sorted(MANIFEST) for k in sorted(MANIFEST[c])]
OUT_FLAGS = {"-A", "-y", "-E", "-a", "-k", "-v", "-u", "-o"}


def replay(binary, case, cname, tmp_path):
    cmd = MANIFEST[case][cname]
    d = os.path.join(GOLD, case)
    stdout_to = None
    if " > " in cmd:
        cmd, stdout_to = cmd.split(" > ")
    toks, args, prev = cmd.split(), [], None
    for x in toks:
        args.append(os.path.join(d, x) if (prev not in OUT_FLAGS and os.path.exists(os.path.join(d, x))) else x)
        prev = x
    out = tmp_path / f"{case}_{cname}"
    out.mkdir()
    with open(out / stdout_to if stdout_to else os.devnull, "wb") as so:
        p = subprocess.run([binary] + args, cwd=out, stdout=so, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    exp_dir = os.path.join(d, "expected", cname)
    for fn in sorted(os.listdir(exp_dir)):
        exp = open(os.path.join(exp_dir, fn), "rb").read()
        if fn.endswith(".bam.raw"):
            got = gzip.open(out / fn[:-4]).read()
        else:
            got = open(out / fn, "rb").read()
        assert got == exp, f"{case}/{cname}/{fn} differs from the reference binary's output"


@pytest.mark.parametrize("case,cname", CASES)
def test_port_matches_reference_golden(case, cname, tmp_path):
    replay(op.PORT_BIN, case, cname, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("case,cname", CASES)
def test_cuda_cli_matches_reference_golden(case, cname, tmp_path):
    from lr2rmats_b200 import api
    assert os.path.exists(api.CLI_PATH), "product CLI not built (run __graft_entry__.build())"
    replay(api.CLI_PATH, case, cname, tmp_path)
