import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
t = symm.empty((1024, 16), dtype=torch.float32, device=f"cuda:{local}")
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
t.fill_(-1.0)
hdl.barrier()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1024, 16), torch.float32)
pt[rank * 4:(rank + 1) * 4].fill_(float(rank + 10))     # write into the peer's memory
torch.cuda.synchronize()
hdl.barrier()
torch.cuda.synchronize()
src = (rank - 1) % world
print(rank, "got from", src, t[src * 4, 0].item(), "untouched", t[1000, 0].item(), flush=True)
dist.destroy_process_group()
