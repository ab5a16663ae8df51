import sys, torch
sys.path.insert(0, '/root/repo')
import bench
print(bench.training_block(torch, torch.device('cuda', 0), with_reference=False))
