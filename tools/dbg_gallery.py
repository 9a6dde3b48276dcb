import sys, torch
sys.path.insert(0, '.')
from peppa_b200 import ops
from peppa_b200.gallery import GalleryStep
from bench import synth_embeddings
dev = torch.device('cuda', 0)
for n in (4096, 65536):
    a, v = synth_embeddings(n, 666, dev)
    ra, na = ops.row_norms(a); rv, nv = ops.row_norms(v)
    print(n, 'ra nan', ra.isnan().sum().item(), ra[:3].tolist(), 'rv nan', rv.isnan().sum().item())
    diag, thr = ops.sim_diag(a, v, ra, rv)
    print('diag nan', diag.isnan().sum().item(), diag[:4].tolist(), 'thr', thr[:4].tolist())
    pd = ops.pair_dot(a, v, rinv_x=ra, rinv_y=rv)
    print('pair_dot', pd[:4].tolist(), (pd - diag).abs().max().item())
    out = GalleryStep(n, 512).run(a, v)
    print('loss', out['loss'].item(), 'recall', out['recall'].tolist()[:3], 'dA nan', out['dA'].isnan().sum().item(), 'ranks', out['ranks'][:8].tolist())
