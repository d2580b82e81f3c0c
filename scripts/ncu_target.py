"""Short target for ncu captures: 1 warm + 1 measured forward step (32 pairs, ViT-L/DPT 512x384) and a few
alignment iterations (config 3).  Usage: ncu ... python scripts/ncu_target.py [forward|align|both] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
what = sys.argv[1] if len(sys.argv) > 1 else 'both'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
if what in ('forward', 'both'):
    from bench import build_model, H, W
    net, cfg = build_model(torch.device('cuda:0'))
    packed = net.repack()
    imgs = torch.rand((2 * B, 3, H, W), device='cuda') * 2 - 1
    idx1, idx2 = np.arange(B, dtype=np.int32), B + np.arange(B, dtype=np.int32)
    for _ in range(2):
        packed.forward(imgs, idx1, idx2, B, H, W)
    torch.cuda.synchronize()
    del packed, net, imgs
    torch.cuda.empty_cache()
if what in ('align', 'both', 'align50'):
    from dust3r_b200.cloud_opt import global_aligner, GlobalAlignerMode
    from scripts.align_config5 import synth_on_device
    n = 50 if what == 'align50' else 8      # align50: BASELINE configs[4] (1225 pairs, ModularPointCloudOptimizer)
    edges = [(i, j) for i in range(n) for j in range(i)]
    out = synth_on_device(n, edges, 384, 512, torch.device('cuda'))
    net = global_aligner(out, 'cuda', mode=GlobalAlignerMode.ModularPointCloudOptimizer if n == 50 else GlobalAlignerMode.PointCloudOptimizer, verbose=False)
    eng = net._get_engine(); net._engine_push(eng)
    eng.run(6)
    torch.cuda.synchronize()
