"""Host <-> device copy bandwidth per rank, alone and with all ranks copying at once (names the limiter of the
end-to-end path at N GPUs).  Launch: python -m torch.distributed.run --nproc-per-node N tools/h2d_bw.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
bind = os.environ.get('BIND', '1') == '1'
if bind:
    bench._bind_to_gpu_numa_node(lr)
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('gloo')
n = 1 << 27                                            # 1 GiB of fp64
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device='cuda'); d_out = torch.ones(n, dtype=torch.float64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n * 8 / dt / 1e9
run(True, True, 2)
res = {}
for name, a, b in (('h2d', True, False), ('d2h', False, True), ('both_each_direction', True, True)):
    res[name + '_all_ranks'] = run(a, b)
# one rank at a time
for name, a, b in (('h2d', True, False), ('d2h', False, True)):
    for r in range(world):
        if world > 1: dist.barrier()
        if r == rank:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(4):
                if a: d_in.copy_(h_in, non_blocking=True)
                if b: h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            res[name + '_alone'] = 4 * n * 8 / (time.perf_counter() - t0) / 1e9
        if world > 1: dist.barrier()
res['cpus'] = len(os.sched_getaffinity(0))
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    agg = {k: round(sum(o[k] for o in out), 1) for k in out[0] if k != 'cpus'}
    print(json.dumps({'world': world, 'numa_bind': bind, 'sum_over_ranks_GBps': agg,
                      'per_rank': [{k: round(v, 1) for k, v in o.items()} for o in out]}), file=sys.stderr)
if world > 1: dist.destroy_process_group()
