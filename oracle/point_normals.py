"""TEST INFRASTRUCTURE ONLY — restatement of MVS::EstimatePointNormals (libs/MVS/DepthMap.cpp:2221-2269). Only tests/ may import this.

CGAL::pca_estimate_normals (third party, absent from /root/reference; its published algorithm): for every point the K nearest
neighbours (the query point included: K+1 points) are fitted with a plane by principal component analysis
(linear_least_squares_fitting_3: centroid + covariance, normal = eigenvector of the smallest eigenvalue); MVS then flips the normal
to face the first view that sees the point (:2262-2265). Here: scipy's cKDTree for the neighbours, numpy eigh (f64) for the PCA —
independent of the product's grid search and Jacobi solver. Ties at the (K+1)-th distance are the one ambiguity (as in CGAL)."""
import numpy as np


def estimate_point_normals(points, view_offsets, views, cam_centers, num_neighbors=16):
    from scipy.spatial import cKDTree
    pts = np.asarray(points, np.float32).astype(np.float64)
    k1 = min(num_neighbors + 1, len(pts))
    _, idx = cKDTree(pts).query(pts, k=k1)
    nb = pts[idx.reshape(len(pts), k1)]                                  # (n, k1, 3)
    c = nb - nb.mean(axis=1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c)
    w, v = np.linalg.eigh(cov)
    nrm = v[:, :, 0]                                                      # eigenvalues ascending
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    first = np.asarray(views)[np.asarray(view_offsets[:-1]).astype(np.int64)]
    C = np.asarray(cam_centers, np.float64).astype(np.float32).astype(np.float64)[first]
    flip = (nrm * (C - pts)).sum(axis=1) < 0
    nrm[flip] *= -1
    gap = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)               # how well the least-variance direction is determined
    return nrm.astype(np.float32), gap
