"""(multi-GPU box, under torchrun) time the 116 MB fp32 gradient all-reduce: NCCL (current env) and, when available,
torch symmetric-memory multimem (NVLS) all-reduce.  Prints one line per variant from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 29_000_090 // 4 * 4
    iters = 40

    def timeit(fn, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {"world": world, "bytes": n * 4, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
    x = torch.ones(n, device=dev)
    ms = timeit(lambda: dist.all_reduce(x))
    out["nccl_ms"] = ms
    out["nccl_busbw_GBs"] = 2 * (world - 1) / world * n * 4 / (ms * 1e-3) / 1e9
    try:
        import torch.distributed._symmetric_memory as symm_mem
        gname = dist.group.WORLD.group_name
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, gname)
        t.fill_(1.0)
        out["multicast"] = bool(getattr(hdl, "multicast_ptr", 0))
        out["zero_symm_ms"] = timeit(lambda: t.zero_())
        out["zero_plain_ms"] = timeit(lambda: x.zero_())
        t.fill_(1.0)
        for opname in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
            op = getattr(torch.ops.symm_mem, opname, None)
            if op is None:
                continue
            try:
                ms2 = timeit(lambda: op(t, "sum", gname))
                out[opname + "_ms"] = ms2
            except Exception as e:  # noqa: BLE001
                out[opname + "_err"] = str(e)[:200]
        t.fill_(1.0); torch.cuda.synchronize(); dist.barrier()
        op = getattr(torch.ops.symm_mem, "multimem_all_reduce_", None) or getattr(torch.ops.symm_mem, "two_shot_all_reduce_")
        op(t, "sum", gname); torch.cuda.synchronize()
        out["symm_result_ok"] = bool((t == world).all())
    except Exception as e:  # noqa: BLE001
        out["symm_mem_error"] = f"{type(e).__name__}: {e}"[:300]
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
