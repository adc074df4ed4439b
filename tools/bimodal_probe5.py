import sys, json, torch
sys.path.insert(0, '/root/repo')
import spectrogram_b200 as sg
import bench
eng = sg.Engine(0)
dev = torch.device('cuda', 0)
stream = torch.cuda.Stream(dev)
c = bench.measure_configs(torch, sg, eng, dev, stream, 6551.4, quick=True)
for n, w in c['config3b_tau'].items():
    print(n, w['kernel'], round(w['ms'], 3), w['steps'])
