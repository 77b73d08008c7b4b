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


def test_two_ranks_shard_clusters_and_rank0_merges(tmp_path):
    """world_size 2 over gloo (CPU): each rank runs the batches of the clusters shard.assign_clusters gives it, rank 0 does
    the cross-batch merge and writes the final files -- byte-identical to the single-process reference run."""
    import socket
    for name in ("in_ref", "in_b200"):
        os.makedirs(tmp_path / name)
    for cl, (seed, n) in enumerate([(21, 30), (22, 16), (23, 22)]):
        reads, _iso, _t = synth.make_cluster(seed=seed, gene_len=500, n_isoforms=2, n_reads=n, eps=0.01)
        synth.write_fastq(str(tmp_path / "in_ref" / ("%d.fastq" % cl)), reads)
        synth.write_fastq(str(tmp_path / "in_b200" / ("%d.fastq" % cl)), reads)
    flags = ["--t", "1", "--split_wrt_batches", "--max_seqs", "20", "--iso_abundance", "2"]
    r0 = subprocess.run([sys.executable, os.path.join(REF, "isONform_parallel"), "--fastq_folder", str(tmp_path / "in_ref"),
                         "--outfolder", str(tmp_path / "out_ref"), "--tmpdir", str(tmp_path / "tmp_ref")] + flags,
                        capture_output=True, text=True, env=_env(), cwd=str(tmp_path), timeout=1500)
    assert r0.returncode == 0, r0.stderr[-3000:]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(_env(), RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "parallel_child.py"), REF, "--fastq_folder",
                                       str(tmp_path / "in_b200"), "--outfolder", str(tmp_path / "out_b200"), "--tmpdir",
                                       str(tmp_path / "tmp_b200")] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                      text=True, env=env, cwd=str(tmp_path)))
    outs = [p.communicate(timeout=1500) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:] + so[-1000:]
    for name in ("transcriptome.fasta", "transcriptome_mapping.txt", "transcriptome_support.txt"):
        assert open(tmp_path / "out_ref" / name).read() == open(tmp_path / "out_b200" / name).read(), name
    batches = [json.loads(re.search(r"B200_PARALLEL (\{.*\})", se).group(1))["batches"] for _so, se in outs]
    assert sum(batches) == 5 and min(batches) >= 1            # 2 + 1 + 2 batches, both ranks worked


def test_in_process_poa_replaces_the_spoa_subprocess(tmp_path):
    """with `--in_process_poa` no consensus spawns a process.  The unmodified reference runs with the
    SAME POA code as its `spoa` executable (tests/stubs_poa/spoa -> scripts/isof_spoa); the final files must be
    byte-identical, and the in-process run must not have called the executable at all."""
    for name in ("in_ref", "in_b200"):
        os.makedirs(tmp_path / name)
    reads, _iso, _t = synth.make_cluster(seed=31, gene_len=500, n_isoforms=2, n_reads=36, eps=0.02)
    synth.write_fastq(str(tmp_path / "in_ref" / "0.fastq"), reads)
    synth.write_fastq(str(tmp_path / "in_b200" / "0.fastq"), reads)
    env = _env()
    env["PATH"] = os.pathsep.join([os.path.join(HERE, "stubs_poa"), env["PATH"]])
    flags = ["--t", "1", "--iso_abundance", "2"]
    r0 = subprocess.run([sys.executable, os.path.join(REF, "isONform_parallel"), "--fastq_folder", str(tmp_path / "in_ref"),
                         "--outfolder", str(tmp_path / "out_ref")] + flags, capture_output=True, text=True, env=env,
                        cwd=str(tmp_path), timeout=1500)
    assert r0.returncode == 0, r0.stderr[-3000:] + r0.stdout[-2000:]
    env_b = dict(env, PATH=os.pathsep.join([os.path.dirname(sys.executable), "/usr/bin", "/bin"]))      # no `spoa` on PATH at all
    r1 = subprocess.run([sys.executable, os.path.join(HERE, "parallel_child.py"), REF, "--fastq_folder", str(tmp_path / "in_b200"),
                         "--outfolder", str(tmp_path / "out_b200"), "--in_process_poa"] + flags, capture_output=True, text=True,
                        env=env_b, cwd=str(tmp_path), timeout=1500)
    assert r1.returncode == 0, r1.stderr[-3000:] + r1.stdout[-2000:]
    for name in ("transcriptome.fasta", "transcriptome_mapping.txt", "transcriptome_support.txt"):
        assert open(tmp_path / "out_ref" / name).read() == open(tmp_path / "out_b200" / name).read(), name
    assert len(open(tmp_path / "out_b200" / "transcriptome.fasta").read()) > 0
