import sys; sys.path.insert(0,'.')
import numpy as np
from srba_stereo_slam_b200 import frontend
rng=np.random.default_rng(0)
img=(rng.random((2,376,1241))*255).astype(np.uint8)
fe=frontend.StereoFrontEnd(max_pairs=1,use_graph=0)
fe.process_stereo(img[0],img[1])
print('ok',[len(a) for a in fe.get_values(0)])
