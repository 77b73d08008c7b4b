"""Child of tests/test_parallel_entry.py: isonform_b200.parallel.InProcessRunner over the oracle-backed test double of
the GPU context (build container only: needs /root/reference)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))

from fake_ctx import FakeContext  # noqa: E402
from isonform_b200 import parallel  # noqa: E402


def main():
    ref = sys.argv[1]
    args = parallel.parser().parse_args(["--reference", ref] + sys.argv[2:])
    ctx = FakeContext()
    barrier = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # world_size-2 test: gloo carries the barriers
        import torch.distributed as dist
        dist.init_process_group("gloo")
        barrier = dist.barrier
    runner = parallel.InProcessRunner(ref, ctx=ctx, barrier=barrier, in_process_poa=args.in_process_poa)
    runner.run(args)
    if barrier is not None:
        import torch.distributed as dist
        dist.destroy_process_group()
    b = runner.memo.bubbles.stats
    sys.stderr.write("B200_PARALLEL " + json.dumps({"batches": runner.batches_run, "calls": ctx.calls, "memo_hits": runner.memo.hits,
                                                    "memo_misses": runner.memo.misses, "bubbles": b}) + "\n")


if __name__ == "__main__":
    main()
