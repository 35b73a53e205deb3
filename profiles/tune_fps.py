"""FPS launch-shape sweep on the three big calls of the eval step (b=32): forces (cluster size, threads) through the
test hooks and times each shape with CUDA events.  Used to calibrate the cost model in csrc/fps.cu."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pu3 = importlib.import_module("3pu_pytorch_b200")
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
dev = torch.device("cuda:0")
for (n, m) in [(6240, 1248), (12480, 2496), (24960, 4992)]:
    x = torch.rand(32, n, 3, device=dev)
    idx = torch.empty(32, m, dtype=torch.int32, device=dev)
    ref = None
    rows = []
    for S in (1, 2, 4, 8):
        for th in (1024, 896, 768, 640, 512, 384, 256, 128):
            lib.pu3_fps_set_cluster(S); lib.pu3_fps_set_threads(th)
            try:
                st = pu3._lib.lib().pu3_fps_f32(32, n, m, x.data_ptr(), None, idx.data_ptr(), torch.cuda.current_stream().cuda_stream)
                if st != 0:
                    continue
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                pu3._lib.lib().pu3_fps_f32(32, n, m, x.data_ptr(), None, idx.data_ptr(), torch.cuda.current_stream().cuda_stream)
                e1.record(); torch.cuda.synchronize()
                if ref is None: ref = idx.clone()
                ok = bool(torch.equal(ref, idx))
                rows.append((e0.elapsed_time(e1), S, th, ok))
            except Exception as e:
                pass
    lib.pu3_fps_set_cluster(0); lib.pu3_fps_set_threads(0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    pu3._lib.lib().pu3_fps_f32(32, n, m, x.data_ptr(), None, idx.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e0.record(); pu3._lib.lib().pu3_fps_f32(32, n, m, x.data_ptr(), None, idx.data_ptr(), torch.cuda.current_stream().cuda_stream); e1.record()
    torch.cuda.synchronize()
    print(f"n={n} m={m}: auto {e0.elapsed_time(e1):.3f} ms; " + "  ".join(f"S{S}x{th}:{ms:.2f}{'' if ok else '!'}" for ms, S, th, ok in sorted(rows)[:8]))
