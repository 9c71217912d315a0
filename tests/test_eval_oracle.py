def test_evaluator_known_answer(golden):
    """plot_utils/kittievalodom.py eval() on the shipped KITTI-03 data."""
    import numpy as np
    from oracle import kitti_eval
    g = golden("kitti03_eval.npz")
    got = kitti_eval.evaluate(g["gt"], g["pred"])
    assert np.allclose(got, g["expected"], rtol=1e-12, atol=0)
    assert np.allclose(got, (11.730234826475684, 0.14661817892056406, 0.16676510581484608, 560.8884529565036), rtol=1e-9)
