import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import uam_path_planning_b200 as uam
H, W, geo = 4099, 4113, (0.0, 32.0 / 4113, 0.0, 32.0 / 4099)
rng = np.random.default_rng(H + W)
m = uam.RegionMap()
for r in ('A', 'B', 'C'):
    m.new_region(r, 'r')
def rect(c, hw, hh, a):
    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    return uam.polygon(*(c + np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T).tolist())
kinds = []
for k in range(260):
    c = rng.uniform(-2, 34, 2)
    t = k % 6 if k % 24 < 6 or k % 6 != 5 else 0
    if t == 0: sh = rect(c, *rng.uniform(0.05, 2.5, 2), rng.uniform(0, np.pi))
    elif t == 1: sh = uam.ball(c.tolist(), float(rng.uniform(0.05, 3.0)), float(rng.uniform(0.02, 3.0)))
    elif t == 2: sh = uam.square(c.tolist(), float(rng.uniform(0.05, 2.0)), float(rng.uniform(0.05, 2.0)))
    elif t == 3: sh = uam.polygon(*(c + rng.uniform(-1.5, 1.5, (3, 2))).tolist())
    elif t == 4: sh = rect(c, rng.uniform(2.0, 9.0), rng.uniform(0.001, 0.01), rng.uniform(0, np.pi))
    else: sh = rect(c, *rng.uniform(4.0, 12.0, 2), rng.uniform(0, np.pi))
    kinds.append(t)
    m.add_obstacle(sh); m.add_shape_to_region('ABC'[k % 3], sh)
eng = m.engine()
for e in (0.0, 0.04, -0.03):
    eng.set_option('rasterizer', 0); a = eng.rasterize_layers(H, W, geo, e)
    eng.set_option('rasterizer', 1); b = eng.rasterize_layers(H, W, geo, e)
    d = (a.view(torch.int32) != b.view(torch.int32))
    n = int(d.sum())
    print('e', e, 'mismatches', n)
    if n:
        idx = torch.nonzero(d)[:12].cpu().numpy()
        for l, i, j in idx:
            print('  layer', l, 'cell', i, j, 'percell', float(a[l, i, j]), 'rows', float(b[l, i, j]), 'sup', i // 256, j // 256, 'col in sup', j % 256)
        li = torch.nonzero(d).cpu().numpy()
        print('  rows with mismatches', len(np.unique(li[:, 1])), 'cols range', li[:, 2].min(), li[:, 2].max(), 'col%256 hist', np.bincount(li[:, 2] % 256, minlength=256).nonzero()[0][:40])
        # which shapes contain the first mismatching cell
        l, i, j = idx[0]
        x, y = geo[0] + (j + 0.5) * geo[1], geo[2] + (i + 0.5) * geo[3]
        for k, sh in enumerate(m.regions['ABC'[l]]['shapes']):
            if sh.contains([x, y]): print('   inside shape', k, sh.kind, sh.records()[0])
