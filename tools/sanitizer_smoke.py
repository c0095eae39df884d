"""Exercises every kernel of the library once (both arithmetic modes, robust 1-to-1, BoW, keyframe database, matchers) at a
small size; run it under `compute-sanitizer --tool memcheck|racecheck` on a B200 (round 1: 0 errors, 0 hazards)."""
import sys; sys.path.insert(0,'.')
import numpy as np
from srba_stereo_slam_b200 import frontend, synth
v=np.load('tests/golden/voc_flat.npz'); voc={k:(v[k] if v[k].ndim else v[k].item()) for k in v.files}
frames=synth.stereo_batch(4,640,192,seed=1,unique=1)
for mode in ((0,0),(1,1)):
    fe=frontend.StereoFrontEnd(width=640,height=192,nlevels=4,nfeatures=300,max_pairs=4,blur_mode=mode[0],pyramid_mode=mode[1],robust_1to1=1,use_graph=0)
    fe.process_batch_host(frames); print([len(a) for a in fe.get_values(3)])
    bow=frontend.BowDatabase(fe,voc); db=frontend.KeyframeDatabase(fe,bow,16)
    db.process_batch(4); r=db.results(4); print(r['query_n'], r['pair_kf'][3])
    fe.process_batch_host(frames[:2]); db.process_batch(2); print(db.size())
    d=fe.get_values(0)[2]; print(len(fe.match_bf(d,d)), bow.query(d,4)[:1])
    db.close(); bow.close(); fe.close()
print('done')
