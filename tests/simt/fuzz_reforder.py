# The reference-order search (warp-cooperative replay, per-voxel truncation, serial replay, counted fallback) on the crowded /
# clustered / duplicate-laden maps of tests/test_gpu_lio.py, under the SIMT emulator: the same test bodies, so that the CPU suite
# exercises the kernels' logic position by position against the oracle.
import test_gpu_lio as G

for per_voxel, nearby in ((8, 18), (14, 18), (3, 74), (9, 6)):
    G.test_reference_order_on_crowded_voxels(per_voxel, nearby)
G.test_reference_order_truncates_crowded_voxels_in_short_sequences()
print("FUZZ_REFORDER_OK")
