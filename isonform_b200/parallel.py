"""In-process drop-in for the reference's two entry points (`isONform_parallel`, per-batch `main`).

    PYTHONHASHSEED=0 python -m isonform_b200.parallel --reference /path/to/isONform \
        --fastq_folder clusters/ --outfolder out/ [--split_wrt_batches ...the isONform_parallel flags]

The reference's driver builds (cluster, batch) instances and hands them to
`multiprocessing.Pool.imap_unordered(isONform, instances)`; the worker `isONform` spawns one
`python main ...` subprocess per batch (isONform_parallel:66-93, 250-311).  With a GPU behind the
seams that layer is the wrong shape: every subprocess would pay an interpreter start, the imports and
a CUDA context, and batches could not share launches.  This module loads both reference scripts as
modules (they carry no `.py` suffix), rebinds the hot-path seams (`patch.install`: minimizers, pair
database / span search, the aligner, the all-pairs merge drivers, the bubble-popping prefetcher) and
replaces exactly two names of the loaded driver module:

    isONform   -> `_run_batch_in_process`: calls the loaded `main.main(args)` directly, stdout/stderr
                  redirected to the same per-batch files the reference writes (:73-77)
    Pool       -> `_RankPool`: runs the instances in this process; under torchrun only those of the
                  clusters `shard.assign_clusters` gives this rank (whole clusters per GPU, so the
                  cross-batch merge of a cluster stays local -- SURVEY.md section 8e)

Everything else -- splitting clusters into batches, instance order, the cross-batch merge and the
final `transcriptome.fasta` / `_mapping.txt` / `_support.txt` -- is the reference's own code, run
unmodified.  `PYTHONHASHSEED=0` must be exported.
"""
import argparse
import contextlib
import importlib.machinery
import importlib.util
import os
import sys
import types


def load_script(path, name):
    """Import an extension-less Python script (the reference's `main`, `isONform_parallel`) as a module
    without running its `__main__` block."""
    loader = importlib.machinery.SourceFileLoader(name, path)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader(name, loader))
    loader.exec_module(mod)
    return mod


def batch_args(fastq, outfolder, params):
    """The namespace `python main ...` would parse from the argv isONform_parallel builds (:79-87) plus the
    defaults of main's own parser (main:585-620) and its xmin clamp (main:623-625)."""
    k = int(params["k"])
    xmin = max(int(params["xmin"]), 2 * k)
    return types.SimpleNamespace(
        fastq=fastq, outfolder=outfolder, k=k, w=int(params["w"]), xmin=xmin, xmax=int(params["xmax"]), T=0.1,
        exact=True, disable_numpy=False, delta_len=int(params["delta_len"]), max_seqs_to_spoa=200, max_seqs=1000,
        exact_instance_limit=int(params["exact_instance_limit"]), set_w_dynamically=False, verbose=False,
        compression=False, delta_iso_len_3=int(params["delta_iso_len_3"]), delta_iso_len_5=int(params["delta_iso_len_5"]),
        parallel=True, slow=False)


class InProcessRunner:
    def __init__(self, reference_dir, ctx=None, rank=None, world_size=None, barrier=None, in_process_poa=False):
        if os.environ.get("PYTHONHASHSEED") != "0":
            raise RuntimeError("export PYTHONHASHSEED=0: the reference's minimizer order is CPython's hash(str) and the GPU "
                               "kernel implements it under that seed (the assignment at main:631 comes too late)")
        self.reference_dir = os.path.abspath(reference_dir)
        if self.reference_dir not in sys.path:
            sys.path.insert(0, self.reference_dir)
        from . import patch
        self.main_mod = load_script(os.path.join(self.reference_dir, "main"), "isonform_reference_main")
        self.driver_mod = load_script(os.path.join(self.reference_dir, "isONform_parallel"), "isonform_reference_parallel")
        self.memo = patch.install(vars(self.main_mod), ctx=ctx, in_process_poa=in_process_poa)
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
        self.barrier = barrier or (lambda: None)
        self.batches_run = 0
        runner = self

        class _RankPool:
            """Stand-in for multiprocessing.Pool inside isONform_parallel.main (:300-311)."""

            def __init__(self, processes=None):
                pass

            def imap_unordered(self, fn, instances):
                from . import shard
                sizes = {}
                for inst in instances:                                  # (location, fastq, outfolder, batch_id, params, cl_id)
                    with open(inst[1]) as fh:
                        sizes[inst[5]] = sizes.get(inst[5], 0) + sum(1 for _ in fh) // 4
                mine = set(shard.assign_clusters(sizes, runner.world)[runner.rank]) if runner.world > 1 else set(sizes)
                for inst in instances:
                    if inst[5] in mine:
                        yield fn(inst)

            def close(self):
                pass

            def terminate(self):
                pass

            def join(self):
                runner.barrier()           # every rank's batch pickles are on disk before rank 0 merges

        self.driver_mod.Pool = _RankPool
        self.driver_mod.isONform = self._run_batch_in_process
        if self.world > 1 and self.rank != 0:
            # the cross-batch merge and the final concatenation run once, on rank 0 (file system is the exchange medium,
            # as in the reference)
            noop = lambda *a, **k: None                                   # noqa: E731
            self.driver_mod.batch_merging_parallel = types.SimpleNamespace(join_back_via_batch_merging=noop)
            self.driver_mod.Parallelization_side_functions = types.SimpleNamespace(
                generate_full_output=noop, remove_folders=noop,
                mkdir_p=self.driver_mod.Parallelization_side_functions.mkdir_p)

    # isONform_parallel:66-93 without the subprocess
    def _run_batch_in_process(self, data):
        _location, fastq, outfolder, batch_id, params, cl_id = data
        self.driver_mod.help_functions.mkdir_p(outfolder)
        args = batch_args(fastq, outfolder, params)
        out_path = os.path.join(outfolder, "stdout{0}_{1}.txt".format(cl_id, batch_id))
        with open(os.path.join(outfolder, "stderr.txt"), "w") as err, open(out_path, "w") as out:
            with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
                self.main_mod.main(args)
        self.batches_run += 1
        return batch_id

    def run_batch(self, fastq, outfolder, **params):
        """One batch through the reference's `main(args)` in this process (defaults = isONform_parallel's)."""
        p = dict(k=20, w=31, xmin=18, xmax=80, delta_len=5, exact_instance_limit=50, delta_iso_len_3=30, delta_iso_len_5=50)
        p.update(params)
        os.makedirs(outfolder, exist_ok=True)
        self.main_mod.main(batch_args(fastq, outfolder, p))
        self.batches_run += 1

    def run(self, args):
        """`isONform_parallel.main(args)` with the worker layer replaced (args: its argparse namespace)."""
        if args.outfolder and not os.path.exists(args.outfolder):
            os.makedirs(args.outfolder, exist_ok=True)
        if self.world > 1 and getattr(args, "tmpdir", None):
            args.tmpdir = os.path.join(args.tmpdir, "rank%d" % self.rank)      # every rank splits the clusters for itself
        if self.world > 1 and self.rank != 0:
            self.barrier()                 # rank 0 restructures the input folder first (:31-55 moves files)
        if self.world > 1 and self.rank == 0:
            self.driver_mod.restructure_isoncorrect_output(args.fastq_folder)
            self.barrier()
        self.driver_mod.main(args)
        if self.world > 1:
            self.barrier()


def parser():
    """isONform_parallel's own flags (:331-367), plus --reference."""
    p = argparse.ArgumentParser(description="isONform_parallel with the hot path on B200, batches run in-process")
    p.add_argument("--reference", required=True, help="checkout of aljpetri/isONform (holds `main`, `isONform_parallel`, modules/)")
    p.add_argument("--fastq_folder", type=str, default=False)
    p.add_argument("--t", dest="nr_cores", type=int, default=8)
    p.add_argument("--k", type=int, default=20)
    p.add_argument("--w", type=int, default=31)
    p.add_argument("--xmin", type=int, default=18)
    p.add_argument("--xmax", type=int, default=80)
    p.add_argument("--exact_instance_limit", type=int, default=50)
    p.add_argument("--keep_old", action="store_true")
    p.add_argument("--set_w_dynamically", action="store_true")
    p.add_argument("--max_seqs", type=int, default=1000)
    p.add_argument("--split_wrt_batches", action="store_true")
    p.add_argument("--outfolder", type=str, default=None)
    p.add_argument("--delta_len", type=int, default=5)
    p.add_argument("--delta", type=float, default=0.1)
    p.add_argument("--max_seqs_to_spoa", type=int, default=200)
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--iso_abundance", type=int, default=5)
    p.add_argument("--delta_iso_len_3", type=int, default=30)
    p.add_argument("--delta_iso_len_5", type=int, default=50)
    p.add_argument("--tmpdir", type=str, default=None)
    p.add_argument("--write_fastq", action="store_true")
    p.add_argument("--in_process_poa", action="store_true",
                   help="consensus by the library's host POA instead of spawning the `spoa` binary (not bit-exact with spoa)")
    return p


def main(argv=None):
    args = parser().parse_args(argv)
    barrier = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
        barrier = dist.barrier
    InProcessRunner(args.reference, barrier=barrier, in_process_poa=args.in_process_poa).run(args)


if __name__ == "__main__":
    main()
