"""CPU, build container only: the UNMODIFIED reference `isONform_parallel` (subprocess per batch, `--t 1`) and
`isonform_b200.parallel` (same flags, batches in-process, hot-path seams rebound over the oracle-backed test double of
the GPU context) must write byte-identical transcriptome.fasta / transcriptome_mapping.txt / transcriptome_support.txt
-- the drop-in claim at the top-level entry point, cross-batch merge included (one cluster is split into batches)."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

from isonform_b200 import synth

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUBS = os.path.join(HERE, "stubs")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "isONform_parallel")), reason="reference tree not present")


def _env():
    env = dict(os.environ, PYTHONHASHSEED="0")
    env["PYTHONPATH"] = os.pathsep.join([STUBS, ROOT, env.get("PYTHONPATH", "")])
    env["PATH"] = os.pathsep.join([STUBS, os.path.dirname(sys.executable), env.get("PATH", "")])
    return env


def test_isonform_parallel_entry_point_identical_outputs(tmp_path):
    for name in ("in_ref", "in_b200"):
        os.makedirs(tmp_path / name)
    for cl, (seed, n) in enumerate([(11, 44), (12, 18)]):
        reads, _iso, _t = synth.make_cluster(seed=seed, gene_len=600, n_isoforms=2, n_reads=n, eps=0.01)
        synth.write_fastq(str(tmp_path / "in_ref" / ("%d.fastq" % cl)), reads)
    for f in os.listdir(tmp_path / "in_ref"):
        shutil.copy(tmp_path / "in_ref" / f, tmp_path / "in_b200" / f)
    flags = ["--t", "1", "--split_wrt_batches", "--max_seqs", "25", "--iso_abundance", "2"]
    r0 = subprocess.run([sys.executable, os.path.join(REF, "isONform_parallel"), "--fastq_folder", str(tmp_path / "in_ref"),
                         "--outfolder", str(tmp_path / "out_ref"), "--tmpdir", str(tmp_path / "tmp_ref")] + flags,
                        capture_output=True, text=True, env=_env(), cwd=str(tmp_path), timeout=1500)
    assert r0.returncode == 0, r0.stderr[-3000:] + r0.stdout[-2000:]
    r1 = subprocess.run([sys.executable, os.path.join(HERE, "parallel_child.py"), REF, "--fastq_folder", str(tmp_path / "in_b200"),
                         "--outfolder", str(tmp_path / "out_b200"), "--tmpdir", str(tmp_path / "tmp_b200")] + flags,
                        capture_output=True, text=True, env=_env(), cwd=str(tmp_path), timeout=1500)
    assert r1.returncode == 0, r1.stderr[-3000:] + r1.stdout[-2000:]
    for name in ("transcriptome.fasta", "transcriptome_mapping.txt", "transcriptome_support.txt"):
        a = open(tmp_path / "out_ref" / name).read()
        b = open(tmp_path / "out_b200" / name).read()
        assert a == b, name
    assert len(open(tmp_path / "out_b200" / "transcriptome.fasta").read()) > 0
    m = re.search(r"B200_PARALLEL (\{.*\})", r1.stderr)
    assert m, r1.stderr[-2000:]
    st = json.loads(m.group(1))
    assert st["batches"] == 3                                  # cluster 0: 25 + 19 reads, cluster 1: 18 reads
    assert st["calls"]["minimizers"] == 3 and st["calls"]["spans"] == 3
    assert st["memo_hits"] > 0
