"""Oracle (test infrastructure): integer skeleton bookkeeping, pure Python.

Restates, independently, what the reference computes at module-construction time:

* edge list with the virtual root edge first      -- skeleton.py:306-315 (get_edges)
* all-pairs edge distance                          -- skeleton.py:364-387 (calc_edge_mat)
* neighbour lists ``dist <= d``                    -- skeleton.py:390-411 (find_neighbor)
* pooling cascade (chains, pair merge, new edges)  -- skeleton.py:160-207 (SkeletonPool.__init__)
* conv channel mask blocks                         -- skeleton.py:34-39, 58-61

Pinned by: the pasted cascade at skeleton.py:464-477 and tests/golden/topology.json
(written by oracle/make_golden.py from the real reference).  Parity: bit-exact.
"""
from collections import deque

SMPL24_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

INF = 100000


def edges_from_parents(parents):
    """Edge e>0 is (parent[e], e); edge 0 is the virtual root edge (0, J)."""
    n = len(parents)
    return [(0, n)] + [(int(parents[i]), i) for i in range(1, n)]


def edge_distance(edges):
    """Shortest path between edges, where two edges are adjacent if they share a joint.

    The reference runs Floyd-Warshall; a BFS per source over the same adjacency gives the
    same integers (unreachable pairs keep the reference's sentinel 100000).  Quirk kept
    bit-exactly: the reference's "direct neighbour" pass also matches an edge with itself
    (skeleton.py:372-380), so the diagonal ends up 1, not 0 -- with d == 0 an edge is
    therefore NOT its own neighbour.
    """
    n = len(edges)
    adj = [[j for j in range(n) if j != i and set(edges[i]) & set(edges[j])] for i in range(n)]
    dist = [[INF] * n for _ in range(n)]
    for s in range(n):
        dist[s][s] = 0
        q = deque([s])
        while q:
            u = q.popleft()
            for v in adj[u]:
                if dist[s][v] == INF:
                    dist[s][v] = dist[s][u] + 1
                    q.append(v)
        dist[s][s] = 1
    return dist


def neighbours(edges, d):
    dist = edge_distance(edges)
    n = len(edges)
    return [[j for j in range(n) if dist[i][j] <= d] for i in range(n)]


def pool_cascade(edges, last_pool=False):
    """Returns (seq_list, pooling_list, new_edges) exactly as SkeletonPool.__init__ builds them."""
    degree = {}
    for a, b in edges:
        degree[a] = degree.get(a, 0) + 1
        degree[b] = degree.get(b, 0) + 1

    chains = []
    # iterative DFS that reproduces the recursion order of find_seq: children are
    # visited in edge-index order, and a chain is cut *before* descending past a branch joint.
    stack = [(0, [])]
    while stack:
        joint, chain = stack.pop()
        if degree.get(joint, 0) > 2 and joint != 0:
            chains.append(chain)
            chain = []
        if degree.get(joint, 0) == 1:
            chains.append(chain)
            continue
        kids = [(edge[1], chain + [idx]) for idx, edge in enumerate(edges) if edge[0] == joint]
        stack.extend(reversed(kids))

    # The recursion appends a finished chain when it is *reached*, so the order above
    # (pre-order) must equal the recursive order.  It does because every append happens
    # on first visit of the joint that terminates the chain.
    pooling, new_edges = [], []
    for chain in chains:
        if last_pool:
            pooling.append(list(chain))
            continue
        rest = list(chain)
        if len(rest) % 2 == 1:
            pooling.append([rest[0]])
            new_edges.append(edges[rest[0]])
            rest = rest[1:]
        for i in range(0, len(rest), 2):
            pooling.append([rest[i], rest[i + 1]])
            new_edges.append([edges[rest[i]][0], edges[rest[i + 1]][1]])
    return chains, pooling, new_edges


def hierarchy(parents=SMPL24_PARENTS, num_layers=4, dist=2):
    """Per-level edges / neighbour lists / pooling lists for the encoder cascade."""
    edges = edges_from_parents(parents)
    levels = []
    for i in range(num_layers):
        nb = neighbours(edges, dist)
        seqs, pooling, new_edges = pool_cascade(edges, last_pool=(i == num_layers - 1))
        levels.append(dict(edges=edges, neighbours=nb, seq_list=seqs, pooling_list=pooling, new_edges=new_edges))
        edges = new_edges
    return levels


def mask_blocks(neigh):
    """(j_out, j_in) pairs whose weight block is unmasked."""
    return [(j, k) for j, nb in enumerate(neigh) for k in nb]
