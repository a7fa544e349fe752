"""Hand-derived known-answer tests for the ML (GNN) operators: <= 6-node graphs whose results are exact rationals,
worked out on paper from the published definitions (PyG 1.7.0 ``add_remaining_self_loops`` / ``gcn_norm`` / ``GCNConv`` /
``GINConv``, torch_scatter ``scatter(reduce='mean')``, ``Batch.from_data_list``'s ``__inc__`` rule).  The ML oracle stays
"parity UNPINNED" (no reference-held vector exists): these KATs pin the restatement -- and the CUDA kernels --
to independent arithmetic, not to the reference.

KAT "gcn": 4 nodes, edges (src -> dst, w) in this order:
    0->1 w 4 | 1->1 w 3 (self loop) | 1->1 w 5 (duplicate self loop) | 2->0 w 3 | 3->3 w 0 (self loop of weight 0)
  add_remaining_self_loops(fill = 1): non-loop edges keep their order, then ONE loop per node; a node with existing loops
  takes the weight of its LAST one (index assignment with duplicates: last write wins):
    0->1 (4), 2->0 (3), 0->0 (1), 1->1 (5), 2->2 (1), 3->3 (0)
  deg (weights summed at the TARGET): deg0 = 3 + 1 = 4, deg1 = 4 + 5 = 9, deg2 = 1, deg3 = 0
  deg^-1/2 = (1/2, 1/3, 1, inf -> 0)
  norm_e = deg^-1/2[src] * w * deg^-1/2[dst]:
    0->1: 1/2 * 4 * 1/3 = 2/3 | 2->0: 1 * 3 * 1/2 = 3/2 | 0->0: 1/4 | 1->1: 5/9 | 2->2: 1 | 3->3: 0
  GCNConv with W = I, b = (1/2, -1/4), X = [[1,2],[3,4],[5,6],[7,8]]:  out_i = sum_{e: dst = i} norm_e * X[src] + b
    out0 = 3/2*(5,6) + 1/4*(1,2)   = (31/4, 19/2)  + b = (33/4, 37/4)
    out1 = 2/3*(1,2) + 5/9*(3,4)   = (7/3, 32/9)   + b = (17/6, 119/36)
    out2 = (5, 6)                                  + b = (11/2, 23/4)
    out3 = 0                                       + b = (1/2, -1/4)
"""
from fractions import Fraction as Fr

GCN_EDGES = [(0, 1, 4.0), (1, 1, 3.0), (1, 1, 5.0), (2, 0, 3.0), (3, 3, 0.0)]
GCN_N = 4
# (src, dst) -> norm after add_remaining_self_loops + gcn_norm
GCN_NORM = {(0, 1): Fr(2, 3), (2, 0): Fr(3, 2), (0, 0): Fr(1, 4), (1, 1): Fr(5, 9), (2, 2): Fr(1), (3, 3): Fr(0)}
GCN_EDGE_ORDER = [(0, 1), (2, 0), (0, 0), (1, 1), (2, 2), (3, 3)]          # PyG order: kept edges, then loops 0..n-1
GCN_X = [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0], [7.0, 8.0]]
GCN_BIAS = [0.5, -0.25]
GCN_OUT = [[Fr(33, 4), Fr(37, 4)], [Fr(17, 6), Fr(119, 36)], [Fr(11, 2), Fr(23, 4)], [Fr(1, 2), Fr(-1, 4)]]
# destination-major CSR of the normalised graph, stable in PyG edge order: row 0 = {2->0, 0->0}, row 1 = {0->1, 1->1}, ...
GCN_CSR_ROWPTR = [0, 2, 4, 5, 6]
GCN_CSR_COL = [2, 0, 0, 1, 2, 3]

# KAT "gin": 3 nodes, x = (1, 10, 100), eps = 1/2, edges 0->1, 0->1 (multi-edge counts twice), 2->1, 1->0
#   pre_i = (1 + eps) * x_i + sum_{j -> i} x_j:   pre0 = 3/2 + 10 = 23/2,  pre1 = 15 + 1 + 1 + 100 = 117,  pre2 = 150
GIN_EDGES = [(0, 1), (0, 1), (2, 1), (1, 0)]
GIN_X = [1.0, 10.0, 100.0]
GIN_EPS = 0.5
GIN_PRE = [Fr(23, 2), Fr(117), Fr(150)]

# KAT "mean": scatter(reduce='mean') with EMPTY segments: x = (2, 4, 9), seg = (0, 0, 2), 4 segments -> (3, 0, 9, 0)
MEAN_X, MEAN_SEG, MEAN_SEGMENTS, MEAN_OUT = [2.0, 4.0, 9.0], [0, 0, 2], 4, [3.0, 0.0, 9.0, 0.0]

# KAT "collate": two samples; request graphs with 3 and 2 nodes, service graph of S = 4 nodes with edges 0->1, 1->2.
#   PyG 1.7.0 __inc__: every attribute whose name contains "index" is offset by the running sum of the samples' num_nodes
#   = REQUEST-graph node counts -> sample 2's service edges are shifted by 3, not by S = 4 (SURVEY 8a-5').
COLLATE_REQ_NODES = [3, 2]
COLLATE_REQ_EDGES = [[(0, 1), (1, 2)], [(0, 1)]]
COLLATE_SVC_EDGES = [(0, 1), (1, 2)]
COLLATE_S = 4
COLLATE_EDGE_INDEX = [[0, 1, 3], [1, 2, 4]]
COLLATE_BATCH = [0, 0, 0, 1, 1]
COLLATE_SVC_FAITHFUL = [[0, 1, 3, 4], [1, 2, 4, 5]]
COLLATE_SVC_SANE = [[0, 1, 4, 5], [1, 2, 5, 6]]
