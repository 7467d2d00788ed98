import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
from seaduck_b200 import parallel
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
n=1_000_000
g=parallel.SnapshotGather(7,n,dev)
print(rank,"mode",g.mode,getattr(g,"peer_error",None),flush=True)
packed=torch.randn(7,n,dtype=torch.float64,device=dev)
for it in range(3):
    g.post(packed,slot=it%2)
g.wait()
# time on the side stream
torch.cuda.synchronize(); dist.barrier()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
t0=time.perf_counter()
with torch.cuda.stream(g.stream):
    e0.record()
for it in range(10):
    g.post(packed,slot=it%2)
with torch.cuda.stream(g.stream):
    e1.record()
host=time.perf_counter()-t0
torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print(rank,"post: device %.3f ms per post (%.1f GB/s out), host issue %.3f ms"%(ms, 56e6*world/ms/1e6, host/10*1e3),flush=True)
# correctness
g.wait()
ref=[torch.empty_like(packed) for _ in range(world)]
dist.all_gather(ref,packed)
ok=all(torch.equal(g.result(1)[r],ref[r]) for r in range(world))
print(rank,"content ok",ok,flush=True)
# raw peer copy timing for one tensor
if world>1 and g.mode=="peer":
    r=(rank+1)%world
    dst=g.peers[r][0][rank]
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): dst.copy_(packed,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(rank,"raw peer copy %.3f ms = %.1f GB/s"%(ms,56e6/ms/1e6),flush=True)
    from seaduck_b200 import _lib
    L=_lib.lib()
    print(rank,"enable",L.sdk_peer_enable(dst.device.index), L.sdk_last_error(),flush=True)
    st=torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): _lib.check(L.sdk_peer_copy_async(dst.data_ptr(), dst.device.index, packed.data_ptr(), local, packed.numel()*8, st),"peer copy")
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(rank,"cudaMemcpyPeerAsync %.3f ms = %.1f GB/s"%(ms,56e6/ms/1e6),flush=True)
    # plain memcpyAsync D2D default kind
    import ctypes
    rt=ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes=[ctypes.c_void_p,ctypes.c_void_p,ctypes.c_size_t,ctypes.c_int,ctypes.c_void_p]
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): rt.cudaMemcpyAsync(dst.data_ptr(), packed.data_ptr(), packed.numel()*8, 4, st)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(rank,"cudaMemcpyAsync(default) %.3f ms = %.1f GB/s"%(ms,56e6/ms/1e6),flush=True)
    # a kernel writing to the peer mapping (SM copy over NVLink)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): dst.view(-1)[:].add_(0.0) if False else None
    e1.record(); torch.cuda.synchronize()
dist.barrier(); dist.destroy_process_group()
