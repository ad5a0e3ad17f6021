"""Host-side helpers of the ctypes harness that involve no device call (CPU)."""
import numpy as np

import pcl_b200 as P


def test_clusters_from_labels_grouping_window_and_order():
    # labels = smallest index of the component, -1 = not clustered (what pclb200_cluster_labels returns)
    labels = np.array([0, 0, 2, -1, 2, 0, 6, 2, 8, 8, 2], dtype=np.int32)
    cl = P.clusters_from_labels(labels)
    assert [c.tolist() for c in cl] == [[2, 4, 7, 10], [0, 1, 5], [8, 9], [6]]  # by size, indices ascending
    assert all(c.dtype == np.int32 for c in cl)
    assert [c.tolist() for c in P.clusters_from_labels(labels, min_size=2, max_size=3)] == [[0, 1, 5], [8, 9]]
    # equal sizes: ordered by the smallest index
    tie = np.array([0, 1, 1, 0, 4, 4], dtype=np.int32)
    assert [c.tolist() for c in P.clusters_from_labels(tie)] == [[0, 3], [1, 2], [4, 5]]
    assert P.clusters_from_labels(np.full(5, -1, np.int32)) == []
    assert P.clusters_from_labels(np.zeros(0, np.int32)) == []


def test_field_and_buffer_views():
    a = np.arange(36, dtype=np.float32).reshape(3, 12)  # three pcl::PointNormal records
    b = P._Buf(a)
    assert (b.rows, b.stride) == (3, 48) and b.ptr.value == a.ctypes.data
    f = P._Buf(P.Field(a, 4))  # the normals inside the same records
    assert (f.rows, f.stride) == (3, 48) and f.ptr.value == a.ctypes.data + 16
    assert P._Buf(None).ptr is None
    assert np.array_equal(P.xyz1(np.ones((2, 3)))[:, 3], [1, 1])
