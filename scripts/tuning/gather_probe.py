"""What an end-of-run object exchange costs under NCCL: dist.all_gather_object against pickle + two tensor all-gathers."""
import os, sys, time, pickle, numpy as np, torch, torch.distributed as dist
rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rng = np.random.default_rng(rank)
def state(n=136):
    return dict(history=(rng.normal(size=(n, 7)), rng.normal(size=n)), seed=5, iter=n, done=True, launches=16,
                weights=rng.normal(size=7).astype(np.float32), init=rng.normal(size=4).astype(np.float32),
                states=[rng.normal(size=4).astype(np.float32) for _ in range(2)], unlucky=None)
obj = [(i, rng.normal(size=7), state()) for i in range(64)]
def manual(o):
    b = pickle.dumps(o, protocol=pickle.HIGHEST_PROTOCOL)
    n = torch.tensor([len(b)], dtype=torch.int64, device="cuda")
    sizes = torch.empty(ws, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(sizes, n)
    sizes = sizes.cpu().tolist()
    m = max(sizes)
    buf = torch.zeros(m, dtype=torch.uint8, device="cuda")
    buf[:len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    out = torch.empty(m * ws, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(out, buf)
    host = out.cpu().numpy()
    return [pickle.loads(host[r * m:r * m + sizes[r]].tobytes()) for r in range(ws)]
for name, fn in (("all_gather_object", lambda o: (lambda out: (dist.all_gather_object(out, o), out)[1])([None] * ws)), ("manual", manual)):
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); r = fn(obj); dt = time.perf_counter() - t0
        if rank == 0: print(name, "rep", rep, "%.2f ms" % (1e3 * dt), len(r), flush=True)
dist.destroy_process_group()
