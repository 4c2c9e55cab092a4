"""Host logic of the spatial slab decomposition (no kernels)."""
import numpy as np


def test_slab_lowering_slices_tables_and_coordinates():
    """Host logic (no kernels): the sub-problem of a slab."""
    from pararealml_b200.operators.fdm.lowering import lower_problem
    from pararealml_b200.operators.fdm.slab import slab_bounds, slab_lowered

    import test_gpu_fused as tf

    cp, _, _ = tf.convection_diffusion_3d_static((23, 19, 38))
    low = lower_problem(cp)
    assert [slab_bounds(23, 3, r) for r in range(3)] == [(0, 8), (8, 16), (16, 23)]
    mid = slab_lowered(low, 6, 18)
    assert mid.shape == (12, 19, 38)
    assert mid.neu_mask & 3 == 0 and mid.dir_mask & 3 == 0
    assert np.array_equal(mid.coords[0], low.coords[0][6:18])
    for f, tab in low.static_dir.items():
        if f >= 2:
            full = tab.reshape(23, -1)
            assert np.array_equal(mid.static_dir[f].reshape(12, -1), full[6:18])
    first = slab_lowered(low, 0, 10)
    assert first.dir_mask & 1 == low.dir_mask & 1 and first.neu_mask & 2 == 0
