import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from helpers import synth, oracle, synth_weights
import gwdepth_b200
from gwdepth_b200 import engine, capi
B, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sd = synth_weights()
images, targets, dgt, sgt = synth.synth_batch(B, H, W, seed=0)
t = time.time(); tr_o = {}
ref = oracle.forward(sd, images, trace=tr_o); print('oracle %.2fs' % (time.time() - t))
eng = engine.Engine(sd)
def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-9)).item(), ((a-b).abs().mean() / b.abs().mean().clamp_min(1e-9)).item()
for mode in ('unpinned', 'pinned'):
    pinned = {}
    if mode == 'pinned':
        pinned = {'line_ids': tr_o['line_ids'].cuda(), 'sample1': tr_o['sample1'].cuda(), 'sample2': tr_o['sample2'].cuda()}
    tr = {}
    out = eng.forward(images.cuda(), pinned=pinned, trace=tr)
    torch.cuda.synchronize()
    print('==', mode, 'launches', capi.launch_count())
    L = tr['c5'].shape[1] * tr['c5'].shape[2]
    print('c5', rel(tr['c5'].permute(0, 3, 1, 2), tr_o['c5']))
    print('memory', rel(tr['memory'].view(B, L, -1).permute(1, 0, 2), tr_o['memory']))
    print('hs', rel(tr['hs'].view(6, B, 100, -1), tr_o['hs']))
    print('logits', rel(out['pred_logits'], ref['pred_logits']), 'lines', rel(out['pred_lines'], ref['pred_lines']))
    print('line ids equal', (tr['line_ids'].cpu().sort(1).values == tr_o['line_ids'].sort(1).values).float().mean().item())
    print('x32', rel(tr['x32'].view(B, L, -1), tr_o['x32']), 'depth0', rel(tr['depth0'].view(B,1,*tr_o['depth0'].shape[-2:]), tr_o['depth0']))
    for i, k in enumerate(['x1', 'x2', 'x3']):
        C = tr_o[k].shape[-1]
        print(k, rel(tr['bufs'][i][:, :C].reshape(B, -1, C), tr_o[k]))
    for k in ('sample1_idx', 'sample2_idx'):
        if k in tr: print(k, 'match frac', (tr[k].cpu().long() == tr_o[k]).float().mean().item())
    for i in range(4):
        print('pred_depth[%d]' % i, rel(out['pred_depth'][i], ref['pred_depth'][i]))
    print('pred_seg', rel(out['pred_seg'], ref['pred_seg']))
