"""2-rank debug of the CUDA-IPC mapping used by P2PSlabSolver (run under torchrun)."""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
import karman_taichi_b200 as kt
from karman_taichi_b200 import _lib
from karman_taichi_b200.slab import P2PSlabSolver, SlabPartition
from karman_taichi_b200.cgsolver import _Handle
lib = _lib.load()
from karman_taichi_b200.slab import _IpcBuffer
f = kt.Field((64, 40), np.float64, dev)
hd = _Handle(f)
buf = _IpcBuffer(hd, 4096)
descs = [None] * world
dist.all_gather_object(descs, buf.handle_bytes())
out = (C.c_double * 2)()
for r in range(world):
    if r == rank: continue
    m = C.c_void_p()
    _lib.check(lib.kt_ipc_open(hd.h, descs[r], C.byref(m)))
    rc = lib.kt_debug_poke(hd.h, m, 7.0 + rank, out)
    print(f'rank {rank}: peer {r} mapped {m.value:#x} poke rc {rc} err {lib.kt_last_error()} before {out[0]} after {out[1]}', flush=True)
dist.barrier(); torch.cuda.synchronize()
rc = lib.kt_debug_poke(hd.h, C.c_void_p(buf.ptr), -1.0, out)
print(f'rank {rank}: my buffer[0] was {out[0]} (written by the peer)', flush=True)
# full solver on a small system
from oracle import krylov
coef, rhs = krylov.poisson_sin_system(64, 0.1)
part = SlabPartition(rhs.shape[0], world)
c_l = np.ascontiguousarray(part.scatter(coef, rank)); b_l = np.ascontiguousarray(part.scatter(rhs, rank))
fc = kt.Field(c_l.shape, np.float64, dev).from_numpy(c_l); fb = kt.Field(b_l.shape, np.float64, dev).from_numpy(b_l)
fx = kt.Field(b_l.shape, np.float64, dev)
s = P2PSlabSolver(fc, fb, fx, part, rank)
print(f'rank {rank}: comm rows {list(s.comm.peer_halo_row)}', flush=True)
try:
    print(f'rank {rank}: solve ->', s.solve(eps=1e-10, max_iter=500), flush=True)
except Exception as e:
    print(f'rank {rank}: solve failed: {e}', flush=True)
dist.destroy_process_group()
