import sys; sys.path.insert(0,'.')
import torch, touchgs_b200 as T
dev=torch.device('cuda:0'); synth=T.synth
cfg=synth.CONFIGS['c3']
sc=synth.make_scene(cfg['N'],3,cfg['smin'],cfg['smax'],0)
cam=synth.orbit_cameras(cfg['W'],cfg['H'],8,3.0,0)[0]
rs=T.GaussianRasterizationSettings(cfg['H'],cfg['W'],cam.tanfovx,cam.tanfovy,torch.zeros(3,device=dev),1.0,cam.viewmatrix.to(dev),cam.projmatrix.to(dev),3,cam.campos.to(dev),False,False)
P=[t.to(dev) for t in (sc.means3D,sc.opacities,sc.shs,sc.scales,sc.rotations)]
ras=T.GaussianRasterizer(rs)
with torch.no_grad():
    for i in range(8): ras(P[0],None,P[1],shs=P[2],scales=P[3],rotations=P[4],tile_rows=(0,9))
torch.cuda.synchronize()
