import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cpu-raytracing-experiments_b200")]
import numpy as np, b2r, scenes, torch
sc = scenes.random_scene(100000)
ps = b2r.PreparedScene(sc, 1920, 1088)
for use_torch_stream in (False, True):
    st = torch.cuda.current_stream().cuda_stream if use_torch_stream else None
    r = b2r.Renderer(ps, 1920, 1088, buckets=8, stream=st)
    print("torch stream" if use_torch_stream else "own stream", st)
    for i in range(3):
        t = time.time(); r.ResetAccumulator(); r.Accumulate(16); r.sync(); print("  Accumulate(16) no upload", time.time() - t)
    for i in range(3):
        t = time.time(); r.SetScene(ps); r.sync(); t1 = time.time() - t
        t = time.time(); r.ResetAccumulator(); r.Accumulate(16); r.sync(); t2 = time.time() - t
        t = time.time(); r.Render(); t3 = time.time() - t
        print("  SetScene %.4f Accumulate(16) %.4f Render %.4f" % (t1, t2, t3))
    r.close()
