"""CPU, build container only: the UNMODIFIED reference `main()` (per-batch entry point, main:341) run on
one synthetic batch twice -- as is, and with `isonform_b200.patch.install` rebinding its hot-path seams
(over the oracle-backed test double of the GPU context) -- must write identical batch pickles
(`<b>_batch`, `spoa<b>`, `mapping<b>`).  This is the drop-in claim at the entry-point level."""
import os
import pickle
import re
import subprocess
import sys

import pytest

from isonform_b200 import synth

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules")), reason="reference tree not present")


def _run(fastq, outdir, patched):
    os.makedirs(outdir, exist_ok=True)
    env = dict(os.environ, PYTHONHASHSEED="0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "ref_e2e_child.py"), "--fastq", fastq, "--outfolder", outdir,
                        "--patched", str(patched)], capture_output=True, text=True, env=env, cwd=outdir, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return r


@pytest.mark.parametrize("seed,n_reads,eps", [(1, 40, 0.01), (5, 30, 0.03)])
def test_reference_main_with_patched_seams_writes_identical_outputs(tmp_path, seed, n_reads, eps):
    reads, _iso, _t = synth.make_cluster(seed=seed, gene_len=700, n_isoforms=3, n_reads=n_reads, eps=eps)
    fastq = str(tmp_path / "7_0.fastq")              # <cluster>_<batch>.fastq as isONform_parallel names it
    synth.write_fastq(fastq, reads)
    r0 = _run(fastq, str(tmp_path / "ref"), 0)
    r = _run(fastq, str(tmp_path / "b200"), 1)
    for name in ("0_batch", "spoa0", "mapping0"):
        a = pickle.load(open(tmp_path / "ref" / name, "rb"))
        b = pickle.load(open(tmp_path / "b200" / name, "rb"))
        assert a == b, name
    assert open(tmp_path / "ref" / "skip0.fa").read() == open(tmp_path / "b200" / "skip0.fa").read()
    spoa = pickle.load(open(tmp_path / "b200" / "spoa0", "rb"))
    assert len(spoa) >= 1
    m = re.search(r"B200_SEAM_CALLS minimizers=(\d+) spans=(\d+) align_calls=(\d+) pairs=(\d+)", r.stderr)
    assert m, r.stderr[-2000:]
    # one minimizer launch and one span-index build for the whole batch; alignments went through the seam
    assert int(m.group(1)) == 1 and int(m.group(2)) == 1 and int(m.group(4)) > 0
    # bubble popping (SimplifyGraph.py:883-1151) is prefetched: one batched alignment call per iteration at most, the
    # bulk of the loop's alignments are memo hits, and speculation never costs an extra spoa spawn
    b = re.search(r"B200_BUBBLES iterations=(\d+) speculated=(\d+) prefetch_calls=(\d+) prefetched=(\d+) spoa_calls=(\d+)",
                  r.stderr)
    hm = re.search(r"memo_hits=(\d+) memo_misses=(\d+)", r.stderr)
    assert b and hm, r.stderr[-2000:]
    iterations, prefetch_calls = int(b.group(1)), int(b.group(3))
    hits, misses = int(hm.group(1)), int(hm.group(2))
    assert iterations >= 1 and prefetch_calls <= iterations
    assert hits > 0 and hits >= misses
    # PathFinder (the fast find_paths) agreed with the reference's own function on every call of the run
    assert int(re.search(r"B200_FIND_PATHS checked=(\d+)", r.stderr).group(1)) > 100
    spawns_ref = int(re.search(r"SPOA_SPAWNS (\d+)", r0.stderr).group(1))
    spawns_b200 = int(re.search(r"SPOA_SPAWNS (\d+)", r.stderr).group(1))
    assert spawns_b200 <= spawns_ref, (spawns_b200, spawns_ref)
