import os, sys, json, subprocess
sys.path.insert(0, '.')
for tiles in (0, 1024):
    for rows in ("0,68", "0,34", "0,9"):
        env = dict(os.environ, TGS_SCATTER_TILES=str(tiles))
        code = f"""
import sys; sys.path.insert(0,'.')
import torch, touchgs_b200 as T
dev=torch.device('cuda:0'); synth=T.synth
cfg=synth.CONFIGS['c3']
sc=synth.make_scene(cfg['N'],3,cfg['smin'],cfg['smax'],0)
cam=synth.orbit_cameras(cfg['W'],cfg['H'],8,3.0,0)[0]
rs=T.GaussianRasterizationSettings(cfg['H'],cfg['W'],cam.tanfovx,cam.tanfovy,torch.zeros(3,device=dev),1.0,cam.viewmatrix.to(dev),cam.projmatrix.to(dev),3,cam.campos.to(dev),False,False)
P=[t.to(dev) for t in (sc.means3D,sc.opacities,sc.shs,sc.scales,sc.rotations)]
r0,r1={rows}
ras=T.GaussianRasterizer(rs)
with torch.no_grad():
    for i in range(5): ras(P[0],None,P[1],shs=P[2],scales=P[3],rotations=P[4],tile_rows=(r0,r1))
    torch.cuda.synchronize(); T._lib.profile_enable(True); T._lib.profile_read()
    for i in range(20): ras(P[0],None,P[1],shs=P[2],scales=P[3],rotations=P[4],tile_rows=(r0,r1))
    torch.cuda.synchronize(); pr=T._lib.profile_read()
print('RES', {tiles}, (r0,r1), ras.last_num_rendered, {{k: round(v[0]/max(v[1],1),4) for k,v in pr.items() if v[1]}})
"""
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print([l for l in r.stdout.splitlines() if l.startswith('RES')] or r.stderr[-500:], flush=True)
