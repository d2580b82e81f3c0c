"""Why is bench.py's cloud_opt section slower than scripts/align_config5.py?  (diagnostic)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device('cuda:0')
pk = bench.peaks()
r = bench.cloud_opt_section(dev, pk)
print('fresh process      :', round(r['ms_per_iter'] * 1e3, 1), 'us/iter; e2e it/s', round(r['e2e']['value']))
r = bench.cloud_opt_section(dev, pk)
print('second call        :', round(r['ms_per_iter'] * 1e3, 1), 'us/iter; e2e it/s', round(r['e2e']['value']))
if 'cpu' in sys.argv:
    bench.cpu_baseline_forward(1)
    r = bench.cloud_opt_section(dev, pk)
    print('after cpu oracle   :', round(r['ms_per_iter'] * 1e3, 1), 'us/iter; e2e it/s', round(r['e2e']['value']))
    torch.set_num_threads(1)
    r = bench.cloud_opt_section(dev, pk)
    print('threads back to 1  :', round(r['ms_per_iter'] * 1e3, 1), 'us/iter; e2e it/s', round(r['e2e']['value']))
