"""(needs a library built with `make -C 3pu_pytorch_b200/csrc -B EXTRA=-DPU3_FPS_TIMELINE`)
In-kernel timeline of the cluster FPS kernel: cycle counter at the phases of rounds 1000..1007 (CTA 0, thread 0).
phases: 0 round start, 1 distance updates done, 2 warp arg-max published, 3 (single-CTA path only) after __syncthreads,
4 the warp's DSMEM stores are issued, 5 mbarrier wait passed, 6 round end."""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
dev = torch.device("cuda:0")
buf = torch.zeros(64, dtype=torch.int32, device=dev)
for (n, m) in [(24960, 4992), (12480, 2496), (6240, 1248)]:
    xyz = torch.rand(32, 3, n, device=dev)
    pu3.operations.furthest_point_sample(xyz, m); torch.cuda.synchronize()
    lib.pu3_fps_set_timeline(ctypes.c_void_p(buf.data_ptr())); buf.zero_()
    pu3.operations.furthest_point_sample(xyz, m); torch.cuda.synchronize()
    lib.pu3_fps_set_timeline(ctypes.c_void_p(0))
    t = (buf.cpu().to(torch.int64) & 0xffffffff).view(8, 8)
    print(f"== n={n} m={m}: cycles within a round (start=0): updates, publish, syncthreads, leader-send, mbar-wait, end; next round start")
    for r in range(7):
        row = [int((t[r, ph] - t[r, 0]) & 0xffffffff) for ph in range(1, 7)]
        print(f"  round {1000 + r}: {row}   period {int((t[r + 1, 0] - t[r, 0]) & 0xffffffff)}")
