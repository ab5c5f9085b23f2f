import os, torch, torch.distributed as dist
rk = int(os.environ["RANK"]); torch.cuda.set_device(rk)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rk}"))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=f"cuda:{rk}")
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rk, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr), "world", h.world_size, "rank", h.rank, flush=True)
    t.zero_(); h.barrier()
    peer = h.get_buffer((rk + 1) % h.world_size, (1024,), torch.uint8)
    peer.fill_(rk + 1)
    h.barrier()
    print(rk, "my buffer now", int(t[0]), flush=True)
    print(rk, "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs][:2], flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
