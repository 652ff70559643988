"""Times the pieces of hal_b200.parallel.all_gather_records under torchrun (diagnostics)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hal_b200 import parallel
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 10_000_000
counts = torch.ones(n, dtype=torch.int64, device="cuda")
recs = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / reps * 1e3
out = torch.empty(world * n * 32, dtype=torch.uint8, device="cuda")
c32 = counts.to(torch.int32); cout = torch.empty(world * n, dtype=torch.int32, device="cuda")
r = {}
r["ag_recs_320MB_into_prealloc"] = t(lambda: dist.all_gather_into_tensor(out, recs))
r["ag_counts_40MB"] = t(lambda: dist.all_gather_into_tensor(cout, c32))
r["alloc_out"] = t(lambda: torch.empty(world * n * 32, dtype=torch.uint8, device="cuda"))
r["cumsum"] = t(lambda: torch.cumsum(cout, 0, dtype=torch.int64))
r["full_all_gather_records"] = t(lambda: parallel.all_gather_records(counts, recs))
# foreign memory (not from torch's allocator), as the library's result buffers are
from cuda.bindings import runtime as cr
err, ptr = cr.cudaMalloc(n * 32)
class _Arr:
    def __init__(self, p, nb): self.__cuda_array_interface__ = {"shape": (nb,), "typestr": "|u1", "data": (int(p), False), "version": 3}
frecs = torch.as_tensor(_Arr(ptr, n * 32), device="cuda")
r["ag_foreign_320MB_direct"] = t(lambda: dist.all_gather_into_tensor(out, frecs))
r["full_foreign_with_copy"] = t(lambda: parallel.all_gather_records(counts, frecs))
if dist.get_rank() == 0:
    print({k: round(v, 2) for k, v in r.items()}, "ms; world", world, flush=True)
dist.destroy_process_group()
