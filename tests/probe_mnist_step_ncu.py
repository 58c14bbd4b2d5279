"""Four eager MNIST steps at B = 65536 (the bench's step without the graph): the target of
`ncu --set full -k regex:"k_skinny|k_fused_static|k_partials|k_reduce_partials" ...` for the small kernels of the step."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
B = 65536
W = ag.workloads
w, x, y = W.mnist_init(np.random.default_rng(1), B)
params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)
dp, state = ag.DataParallel(0, 1), {}
for _ in range(int(os.environ.get("STEPS", 4))):
    W.mnist_train_step(ag, dp, params, xd, yd, 0.1, state)
ag.sync()
