"""Small program for compute-sanitizer: every kernel of the library once, on tiny inputs.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from eigen_value_b200 import EigenValue, Solver  # noqa: E402


def main():
    s = Solver(0)
    for dim in (96, 512):
        d = s.hilbert(dim)
        u = s.uniform(dim, 7)
        for kernel, form in ((0, 0), (1, 0), (2, 0), (1, 1)):
            info, vec = s.solve_device(d, dim, kernel=kernel, form=form, max_iter=6)
            print(dim, "kernel", info.kernel_id, "form", form, "rounds", info.iter_count, float(info.eigen_val))
        info, _ = s.solve_device(u, dim, max_iter=4)
        d.free()
        u.free()
    # multi-unit rows of the resident-e kernel (N > 8192) and the chunked general loop
    d = s.hilbert(8192 + 1024)
    for kernel in (0, 1):
        info, _ = s.solve_device(d, 8192 + 1024, kernel=kernel, max_iter=2)
        print("9216 kernel", info.kernel_id, float(info.eigen_val))
    d.free()
    m = np.arange(1, 65, dtype=np.float32)
    print(s.find_max(m), s.stop(m), s.sum_across_rows(np.eye(33, dtype=np.float32)).sum())
    w = np.ones((7, 7), np.float32)
    s.compute_next_matrix(w, np.full(7, 7.0, np.float32))
    ev = EigenValue()
    print(ev.similarity_transform(np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)))
    print("SANITIZE_TARGET_DONE")


if __name__ == "__main__":
    main()
