"""Batched k-NN probe: builds the bench map and runs lsd_knn_query_dev once per query order.
Meant to be run under ncu (`-k regex:knn_query_kernel`); prints CUDA-event timings otherwise."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import lsdreg  # noqa: E402
from lsdreg import synth  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
lsdreg.init(0)
dev = torch.device("cuda", 0)
m = synth.block_map(bench.MAP_SEED, bench.BLOCKS_X, bench.BLOCKS_Y, bench.SPACING)
hmap = lsdreg.HashVoxelMap(0.5, 25)
hmap.insert(m, 0)
print(json.dumps(bench.run_knn_batch(torch, hmap, m, dev, nq)))
