import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
t0 = time.time()
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(4, device="cuda") * (local + 1)
dist.all_reduce(x); torch.cuda.synchronize()
dist.barrier(); torch.cuda.synchronize()
print(f"rank {dist.get_rank()} ok {x.tolist()} in {time.time()-t0:.1f}s", flush=True)
dist.destroy_process_group()
