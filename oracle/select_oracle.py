"""CPU ORACLE of phyloselect.py's front half -- TEST INFRASTRUCTURE ONLY (never imported by
phyloligo_b200/; tests/ and the golden generator only).

numpy restatement of the K-medoids (PAM) loop of the reference,
phylopackage/bin/phyloselect.py: ``fit`` :119-171, ``_get_cluster_ics`` :187-195,
``_update_medoid_ics_in_place`` :197-240, ``_get_initial_medoid_indices`` :291-309, on a precomputed
distance matrix.  Pinned: tests/golden/select_golden.npz holds the labels, medoids and iteration counts
the reference's OWN class produced (tests/golden/make_select_golden.py executes the class body taken
from /root/reference by AST); tests/test_select_oracle.py replays them here.

The nearest-neighbour graph has no code in the reference: TSNE(metric="precomputed") and
HDBSCAN(metric="precomputed") (:381-428) compute it inside scikit-learn / hdbscan.  Its oracle is
scikit-learn's own routine, NearestNeighbors(metric="precomputed").kneighbors_graph(mode="distance")
(scikit-learn is installed; the pinned version of the reference is 0.19.1, meta.yaml:20).
"""
import warnings

import numpy as np


def initial_medoids(D, n_clusters, init="heuristic", random_state=None):
    """phyloselect.py:291-309"""
    if init == "random":
        rs = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        return rs.permutation(D.shape[0])[:n_clusters]
    if init == "heuristic":
        # the K points with the smallest sum of distances to every other point
        return list(np.argsort(np.sum(D, axis=1))[:n_clusters])
    raise ValueError("Initialization not implemented for method: '{}'".format(init))


def kmedoids_fit(D, n_clusters=8, init="heuristic", max_iter=300, random_state=None):
    """Returns (labels, medoid_indices, n_iter) as KMedoids(distance_metric="precomputed").fit(D)
    leaves them in labels_, (the indices behind) cluster_centers_, n_iter_  (phyloselect.py:119-171)."""
    D = np.asarray(D)
    medoid_ics = initial_medoids(D, n_clusters, init, random_state)
    old_medoid_ics = np.zeros((n_clusters,))
    n_iter = 0
    cluster_ics = None
    while not np.all(old_medoid_ics == medoid_ics) and n_iter < max_iter:
        n_iter += 1
        old_medoid_ics = np.copy(medoid_ics)
        cluster_ics = np.argmin(D[medoid_ics, :], axis=0)                      # :187-195
        for c in range(n_clusters):                                            # :197-240
            members = cluster_ics == c
            if members.sum() == 0:
                warnings.warn("Cluster {} is empty!".format(c))
                continue
            curr_cost = np.sum(D[medoid_ics[c], members])
            all_costs = np.sum(D[members, :][:, members], axis=1)
            best = np.argmin(all_costs)
            if all_costs[best] < curr_cost:
                medoid_ics[c] = np.where(members)[0][best]
    return cluster_ics, np.asarray(medoid_ics, dtype=np.int64), n_iter


def knn_graph(D, k):
    """(indices [n, k], distances [n, k]) of the k nearest neighbours of every row of a precomputed
    matrix, the row itself excluded, ascending -- scikit-learn's kneighbors_graph(mode="distance")."""
    from sklearn.neighbors import NearestNeighbors
    nn = NearestNeighbors(n_neighbors=k, metric="precomputed").fit(D)
    dist, idx = nn.kneighbors()  # no argument: the training points, each one's own entry left out
    return idx, dist
