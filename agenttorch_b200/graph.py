"""Contact graph in the layout the sm_100a step kernel reads: destination-sorted CSR.

Replaces the `(edge_index, edge_attr)` COO pair that the reference hands to PyG
(`network_from_file`, models/covid/substeps/utils.py:129-148; consumed by
`NewTransmission.forward`, new_transmission/transition.py:165-218).

* rows = destination agents (`edge_index[1]`), columns = source agents (`edge_index[0]`);
* the sort is STABLE, so inside a row the edges keep the order of the original edge list —
  the order in which the reference's scatter-add (index_add on CPU) accumulates them;
* edges are used as written: directed, duplicates and self-loops kept;
* `col_idx` is int32, padded by >= 4 entries and 16-byte aligned for 128-bit loads;
* `i_bar` reproduces quirk Q1: the reference's `torch.gather(x_i[:,4], 0, edge_type)` with
  edge_type == c for every edge reads ONE scalar, `mean_interactions[dst[c]]` (c = 1 in
  `network_from_file`), indexed in the ORIGINAL edge order.
"""
from dataclasses import dataclass
from typing import Optional

import torch

COL_PAD = 8


@dataclass
class ContactGraph:
    n_rows: int                      # destination agents owned by this shard
    n_sources: int                   # size of the source id space (code / bitmap arrays)
    row_offset: int                  # global id of local row 0 (multiple of 32)
    row_ptr: torch.Tensor            # (n_rows+1,) int32
    col_idx: torch.Tensor            # (E + COL_PAD,) int32, source ids
    num_edges: int
    edge_weight: Optional[torch.Tensor] = None   # (E + COL_PAD,) float32 or None when all == 1
    i_bar: float = 1.0
    i_bar_agent: int = -1            # agent whose mean_interactions is I_bar (dst of original edge #c)
    # multi-layer contact networks (per-edge network type, transition.py:71-73): layer number of every CSR edge and the I_bar
    # of every layer (mean_interactions[dst[type]] of the layer's type); None / () when every edge has the same type
    edge_layer: Optional[torch.Tensor] = None    # (E + COL_PAD,) uint8
    i_bar_layers: tuple = ()

    @property
    def device(self):
        return self.row_ptr.device

    def to(self, device):
        return ContactGraph(self.n_rows, self.n_sources, self.row_offset, self.row_ptr.to(device),
                            self.col_idx.to(device), self.num_edges,
                            None if self.edge_weight is None else self.edge_weight.to(device),
                            self.i_bar, self.i_bar_agent,
                            None if self.edge_layer is None else self.edge_layer.to(device), tuple(self.i_bar_layers))

    def edge_index(self):
        """COO view in CSR order: (2,E) int64 [src; dst] with global destination ids."""
        deg = (self.row_ptr[1:] - self.row_ptr[:-1]).long()
        dst = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), deg) + self.row_offset
        return torch.stack((self.col_idx[: self.num_edges].long(), dst))

    def avg_degree(self):
        return self.num_edges / max(self.n_rows, 1)


def _pad_cols(col, pad_value=0):
    out = torch.full((col.numel() + COL_PAD,), pad_value, dtype=col.dtype, device=col.device)
    out[: col.numel()] = col
    return out


def build_contact_graph(edge_index, num_agents, edge_attr=None, mean_interactions=None, device=None):
    """COO (reference format) -> ContactGraph.

    edge_index: (2,E) integer tensor [src; dst]; edge_attr: optional (2,E) [edge_type; weight]
    (utils.py:143-145); mean_interactions: (N,) or (N,1) used for I_bar (Q1)."""
    edge_index = torch.as_tensor(edge_index)
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError("edge_index must have shape (2, E)")
    device = torch.device(device) if device is not None else edge_index.device
    e = int(edge_index.shape[1])
    if e >= 2**31 - COL_PAD:
        raise ValueError("more than 2^31 edges in one shard: partition the graph first")
    src = edge_index[0].to(device=device, dtype=torch.int64)
    dst = edge_index[1].to(device=device, dtype=torch.int64)
    if e and (int(src.min()) < 0 or int(src.max()) >= num_agents or int(dst.min()) < 0 or int(dst.max()) >= num_agents):
        raise ValueError("edge endpoint out of range")

    weight = None
    i_bar, i_bar_agent = 1.0, -1
    etype_const = 1
    layer, i_bar_layers = None, ()
    if edge_attr is not None:
        edge_attr = torch.as_tensor(edge_attr)
        etype = edge_attr[0].to(device=device, dtype=torch.int64)
        if e:
            etype_const = int(etype[0])
            if not bool((etype == etype_const).all()):
                # multi-layer network: the reference gathers I_bar per edge, I_bar_e = mean_interactions[dst[type_e]]
                # (transition.py:71-73: a gather of the per-edge vector at index type_e) — one value per distinct type
                types, layer = torch.unique(etype, return_inverse=True)
                if types.numel() > 8:
                    raise NotImplementedError("more than 8 distinct edge types (network layers)")
                if int(types.min()) < 0 or int(types.max()) >= e:
                    raise IndexError("edge_type indexes past the edge list (the reference's gather would fail)")
                if mean_interactions is None:
                    raise ValueError("a multi-layer contact network needs mean_interactions")
                mi = torch.as_tensor(mean_interactions).reshape(-1).to(device)
                i_bar_layers = tuple(float(mi[int(dst[int(k)])]) for k in types)
        w = edge_attr[1].to(device=device, dtype=torch.float32)
        if e and not bool((w == 1.0).all()):
            weight = w
    if weight is not None and layer is not None:
        raise NotImplementedError("edge weights on a multi-layer contact network are not supported")
    if e and mean_interactions is not None:
        if not 0 <= etype_const < e:
            raise IndexError("edge_type indexes past the edge list (the reference's gather would fail)")
        i_bar_agent = int(dst[etype_const])
        i_bar = float(torch.as_tensor(mean_interactions).reshape(-1)[i_bar_agent])

    order = torch.sort(dst, stable=True).indices
    col = src[order].to(torch.int32)
    counts = torch.bincount(dst, minlength=num_agents)
    row_ptr = torch.zeros(num_agents + 1, dtype=torch.int64, device=device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    g = ContactGraph(num_agents, num_agents, 0, row_ptr.to(torch.int32), _pad_cols(col), e,
                     None if weight is None else _pad_cols(weight[order]), i_bar, i_bar_agent,
                     None if layer is None else _pad_cols(layer[order].to(torch.uint8)), i_bar_layers)
    return g


GRAPH_FILE_VERSION = 1


def save_contact_graph(graph, path):
    """Binary CSR on disk: what `build_contact_graph` produced (sorted rows, int32 columns, I_bar), so that a population's
    edge-list CSV (`mobility_networks/0.csv`, 2.5 GB of text for 168 M edges) is parsed and sorted once."""
    if graph.edge_layer is not None:
        raise NotImplementedError("multi-layer contact networks are not stored in the binary graph file")
    cpu = graph.to("cpu")
    torch.save({"version": GRAPH_FILE_VERSION, "n_rows": cpu.n_rows, "n_sources": cpu.n_sources, "row_offset": cpu.row_offset,
                "row_ptr": cpu.row_ptr, "col_idx": cpu.col_idx[: cpu.num_edges].clone(), "num_edges": cpu.num_edges,
                "edge_weight": None if cpu.edge_weight is None else cpu.edge_weight[: cpu.num_edges].clone(),
                "i_bar": cpu.i_bar, "i_bar_agent": cpu.i_bar_agent}, path)


def load_contact_graph(path, device="cpu"):
    d = torch.load(path, map_location="cpu", weights_only=True)
    if d.get("version") != GRAPH_FILE_VERSION:
        raise ValueError(f"{path}: unsupported contact-graph file version {d.get('version')}")
    if d["row_ptr"].numel() != d["n_rows"] + 1 or int(d["row_ptr"][-1]) != d["num_edges"] or d["col_idx"].numel() != d["num_edges"]:
        raise ValueError(f"{path}: inconsistent contact-graph file")
    g = ContactGraph(d["n_rows"], d["n_sources"], d["row_offset"], d["row_ptr"], _pad_cols(d["col_idx"]), d["num_edges"],
                     None if d["edge_weight"] is None else _pad_cols(d["edge_weight"]), d["i_bar"], d["i_bar_agent"])
    return g.to(device)


# Astoria's age histogram (populations/astoria/age.pickle), used for synthetic populations
ASTORIA_AGE_COUNTS = (6723, 11573, 5724, 4647, 3910, 4941)


def synthetic_population(num_agents, seed, initial_exposed=0.01, device="cpu"):
    """age (uint8, categorical with astoria's proportions) and initial stage (float32: 1 for a
    fraction `initial_exposed`, else 0) — the inputs of BASELINE configs C2/C3/C5."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    probs = torch.tensor(ASTORIA_AGE_COUNTS, dtype=torch.float64)
    age = torch.multinomial(probs / probs.sum(), num_agents, replacement=True, generator=gen).to(torch.uint8)
    stage0 = (torch.rand(num_agents, generator=gen) < initial_exposed).to(torch.float32)
    return age.to(device), stage0.to(device)


def mean_interactions_from_age(age):
    """2 / 3 / 4 contacts for children / adults / elderly (utils.py:88-109)."""
    age = age.reshape(-1).to(torch.int64)
    out = torch.full(age.shape, 3.0, dtype=torch.float32, device=age.device)
    out[age < 1] = 2.0
    out[age > 4] = 4.0
    return out


def synthetic_contact_graph(num_agents, avg_degree, seed, family="rand", degree_dist="poisson",
                            block=65536, p_local=0.8, device="cpu", row_begin=0, row_end=None,
                            mean_interactions=None):
    """Seeded synthetic contact graph generated directly in CSR form (no sort needed).

    In-degree ~ Poisson(avg_degree) or negative-binomial with the same mean (dispersion 4, as the
    census generator data/census/generate/mobility_network.py:12-19 does per agent).  Sources:
      family="rand": uniform over all agents (the reference generator's configuration model);
      family="comm": with probability p_local uniform inside the destination's block of `block`
                     consecutive ids, else uniform over all agents (community structure).
    Only rows [row_begin, row_end) are generated (a rank's shard); every row has its own generator
    stream, so shards of different world sizes describe the same global graph when generated in
    chunks aligned to `gen_chunk`.
    """
    row_end = num_agents if row_end is None else row_end
    gen_chunk = 1 << 20
    dev = torch.device(device)
    cols, degs = [], []
    for c0 in range((row_begin // gen_chunk) * gen_chunk, row_end, gen_chunk):
        c1 = min(c0 + gen_chunk, num_agents)
        g = torch.Generator(device=dev).manual_seed(int(seed) * 1000003 + c0 // gen_chunk)
        n = c1 - c0
        if degree_dist == "poisson":
            deg = torch.poisson(torch.full((n,), float(avg_degree), device=dev), generator=g)
        elif degree_dist == "nb":
            k = 4.0
            u = torch.rand((n, int(k)), device=dev, generator=g).clamp_min(1e-12)
            rate = -torch.log(u).sum(1) * (float(avg_degree) / k)          # Gamma(k, d/k)
            deg = torch.poisson(rate, generator=g)
        else:
            raise ValueError(degree_dist)
        deg = deg.to(torch.int64)
        total = int(deg.sum())
        dst_local = torch.repeat_interleave(torch.arange(n, device=dev), deg)
        src = torch.randint(0, num_agents, (total,), device=dev, generator=g)
        if family == "comm":
            local = torch.rand(total, device=dev, generator=g) < p_local
            blk0 = ((dst_local + c0) // block) * block
            blk_n = torch.clamp(torch.full_like(blk0, block), max=num_agents - blk0)
            src_local = blk0 + (torch.rand(total, device=dev, generator=g) * blk_n).long().clamp_max(blk_n - 1)
            src = torch.where(local, src_local, src)
        elif family != "rand":
            raise ValueError(family)
        # keep only the rows of this shard
        lo, hi = max(row_begin, c0) - c0, min(row_end, c1) - c0
        if lo > 0 or hi < n:
            starts = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            starts[1:] = torch.cumsum(deg, 0)
            src = src[int(starts[lo]): int(starts[hi])]
            deg = deg[lo:hi]
        cols.append(src.to(torch.int32))
        degs.append(deg)
    deg = torch.cat(degs) if degs else torch.zeros(0, dtype=torch.int64, device=dev)
    col = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.int32, device=dev)
    e = int(col.numel())
    if e >= 2**31 - COL_PAD:
        raise ValueError("more than 2^31 edges in one shard")
    row_ptr = torch.zeros(row_end - row_begin + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(deg, 0)
    graph = ContactGraph(row_end - row_begin, num_agents, row_begin, row_ptr.to(torch.int32), _pad_cols(col), e)
    if mean_interactions is not None and e > 1 and row_begin == 0:
        # canonical edge list = CSR order; original edge #1 is the second CSR entry (Q1)
        first_rows = torch.searchsorted(row_ptr, torch.tensor([1], device=dev), right=True) - 1
        graph.i_bar_agent = int(first_rows[0])
        graph.i_bar = float(mean_interactions.reshape(-1)[graph.i_bar_agent])
    return graph


def shard_bounds(num_agents, world_size, align=1024):
    """Contiguous ownership ranges, boundaries aligned so that bitmap words never straddle ranks."""
    per = -(-num_agents // world_size)
    per = -(-per // align) * align
    return [(min(r * per, num_agents), min((r + 1) * per, num_agents)) for r in range(world_size)]


def shard_code_len(num_agents, world_size, align=1024):
    """Length of the code / hot arrays for `world_size` equal slices (the last rank's slice is zero-padded)."""
    per = -(-num_agents // world_size)
    per = -(-per // align) * align
    return per * world_size


def shard_graph(graph, row_begin, row_end):
    """Rows [row_begin,row_end) of a full (row_offset == 0) graph as a shard with global column ids."""
    if graph.row_offset != 0:
        raise ValueError("shard_graph expects the full graph")
    rp = graph.row_ptr.long()
    e0, e1 = int(rp[row_begin]), int(rp[row_end])
    col = graph.col_idx[e0:e1].clone()
    w = None if graph.edge_weight is None else _pad_cols(graph.edge_weight[e0:e1].clone())
    lay = None if graph.edge_layer is None else _pad_cols(graph.edge_layer[e0:e1].clone())
    return ContactGraph(row_end - row_begin, graph.n_sources, row_begin, (rp[row_begin: row_end + 1] - e0).to(torch.int32),
                        _pad_cols(col), e1 - e0, w, graph.i_bar, graph.i_bar_agent, lay, tuple(graph.i_bar_layers))


# --------------------------------------------------------------------------------------- BCSR
TAIL_PAD = 2048          # ints of padding after the last segment (TMA chunks may read past a segment's end)


@dataclass
class BlockedGraph:
    """Column-blocked, destination-sorted CSR — the layout the sm_100a kernels stream (include/at_seirm.h).

    Rows are grouped in gather tasks of `rows_per_task` rows (32 with one column block, 256 with several);
    sources are split in `n_blocks` ranges of `block_size` ids so that one block's "hot" bitmap fits in shared
    memory.  Segment s = block*n_gtasks + gtask stores the edges of the task's rows whose source is in the
    block, row by row, each row in original edge order, 16-byte aligned.  `col` holds block-local source ids
    encoded as ((c >> 5) << 7) | (c & 31)."""
    n_rows: int
    n_rows_pad: int
    n_sources: int
    row_offset: int
    n_blocks: int
    block_size: int
    rows_per_task: int
    row_start: torch.Tensor      # (n_blocks, n_rows_pad) int16 holding uint16 offsets inside the segment
    seg_info: torch.Tensor       # (n_blocks*n_gtasks, 2) int32: (first entry / 4, edges); stream layout: (units,) tile of each unit
    col: torch.Tensor            # (padded edges + TAIL_PAD,) int32
    num_edges: int
    edges_per_chunk: int = 768
    hash_words: int = 0          # > 0: col holds raw source ids and the smem bitmap is a hashed filter of that many words
    edge_weight: Optional[torch.Tensor] = None
    i_bar: float = 1.0
    i_bar_agent: int = -1
    layout: int = 0              # 0 = BCSR (row_start + col), 1 = stream (self-describing entries; include/at_seirm.h)
    tile_rows: int = 0           # stream layout: rows per tile
    block_units: Optional[torch.Tensor] = None   # stream layout: (n_blocks+1,) int32 first unit of every source block
    tile_rec_base: Optional[torch.Tensor] = None # stream layout: (n_tiles,) int64 first hit-record slot of every tile
    lam_deg_bits: int = 0        # stream layout: bits of the largest in-degree (head room of the fixed-point sums)
    i_bar_layers: tuple = ()     # stream layout: I_bar of every network layer (multi-layer contact networks), else ()

    @property
    def device(self):
        return self.col.device

    @property
    def n_tasks(self):
        """32-row tasks (= words of the row bitmaps)"""
        return self.n_rows_pad // 32

    def avg_degree(self):
        return self.num_edges / max(self.n_rows, 1)

    def to(self, device):
        return BlockedGraph(self.n_rows, self.n_rows_pad, self.n_sources, self.row_offset, self.n_blocks, self.block_size,
                            self.rows_per_task, self.row_start.to(device), self.seg_info.to(device), self.col.to(device),
                            self.num_edges, self.edges_per_chunk, self.hash_words,
                            None if self.edge_weight is None else self.edge_weight.to(device), self.i_bar, self.i_bar_agent,
                            self.layout, self.tile_rows,
                            None if self.block_units is None else self.block_units.to(device),
                            None if self.tile_rec_base is None else self.tile_rec_base.to(device), self.lam_deg_bits,
                            tuple(self.i_bar_layers))


def choose_block_size(n_sources, max_block_size):
    """Smallest number of equal blocks (multiples of 1024 ids) that fit the shared-memory bitmap capacity."""
    if max_block_size < 1024:
        raise RuntimeError("shared-memory bitmap capacity too small")
    n_blocks = max(1, -(-n_sources // max_block_size))
    block = -(-n_sources // n_blocks)
    block = -(-block // 1024) * 1024
    return n_blocks, block


def choose_edges_per_chunk(avg_edges_per_task):
    """TMA chunk capacity: the smallest of 512/768/1024 edges that holds a typical gather task in one chunk."""
    for cap in (512, 768, 1024):
        if avg_edges_per_task * 1.12 + 16 <= cap:
            return cap
    return 1024


def _to_i16(x):
    x = x.to(torch.int32) & 0xFFFF
    return torch.where(x >= 32768, x - 65536, x).to(torch.int16).contiguous()


def block_graph(graph: ContactGraph, block_size: int, n_blocks: Optional[int] = None, edges_per_chunk: Optional[int] = None,
                hash_words: int = 0) -> BlockedGraph:
    """ContactGraph (plain dst-sorted CSR, global source ids) -> BlockedGraph, on the graph's device.
    hash_words > 0: single block, raw source ids in `col`, hashed shared-memory filter (see include/at_seirm.h)."""
    dev = graph.device
    if graph.edge_layer is not None:
        raise NotImplementedError("multi-layer contact networks need the stream layout")
    n_rows, e = graph.n_rows, graph.num_edges
    n_rows_pad = -(-max(n_rows, 1) // 256) * 256
    if block_size % 1024:
        raise ValueError("block_size must be a multiple of 1024")
    nb = max(1, -(-graph.n_sources // block_size)) if n_blocks is None else n_blocks
    if nb * block_size < graph.n_sources:
        raise ValueError("blocks do not cover the source id space")
    if nb > 64:
        raise ValueError("more than 64 column blocks: shard the graph over more GPUs")
    rpt = 32 if nb == 1 else 256
    shift = 5 if nb == 1 else 8
    n_gt = n_rows_pad // rpt
    row_ptr = graph.row_ptr.long()
    degf = row_ptr[1:] - row_ptr[:-1]
    col = graph.col_idx[:e].long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=dev), degf)
    order = None
    if nb == 1:
        key = rows
    else:
        blk = torch.div(col, block_size, rounding_mode="floor")
        key = blk * n_rows_pad + rows
        order = torch.sort(key, stable=True).indices
        key = key[order]
        col = col[order] - blk[order] * block_size
        del blk
    del rows
    n_seg = nb * n_gt
    deg2 = torch.bincount(key, minlength=nb * n_rows_pad)
    seg = key >> shift
    seg_len = deg2.reshape(n_seg, rpt).sum(1)
    if e and int(seg_len.max()) > 65535:
        raise ValueError("a gather task has more than 65535 edges inside one column block (hub rows are not supported)")
    # offset of each row inside its segment (exclusive prefix of the degrees within the task)
    d3 = deg2.reshape(n_seg, rpt)
    row_start = (torch.cumsum(d3, 1) - d3).reshape(nb, n_rows_pad)
    seg_pad = ((seg_len + 3) >> 2) << 2
    seg_base = torch.zeros(n_seg + 1, dtype=torch.int64, device=dev)
    seg_base[1:] = torch.cumsum(seg_pad, 0)
    seg_first = torch.zeros(n_seg, dtype=torch.int64, device=dev)
    seg_first[1:] = torch.cumsum(seg_len, 0)[:-1]
    total = int(seg_base[-1])
    if total // 4 >= 2**31:
        raise ValueError("graph shard too large for int32 segment bases")
    pos = torch.arange(e, device=dev) - seg_first[seg] + seg_base[:-1][seg]
    enc = col.to(torch.int32) if hash_words else (((col >> 5) << 7) | (col & 31)).to(torch.int32)
    col_out = torch.zeros(total + TAIL_PAD, dtype=torch.int32, device=dev)
    col_out[pos] = enc
    w_out = None
    if graph.edge_weight is not None:
        w = graph.edge_weight[:e]
        w_out = torch.ones(total + TAIL_PAD, dtype=torch.float32, device=dev)
        w_out[pos] = w if order is None else w[order]
    seg_info = torch.stack(((seg_base[:-1] >> 2), seg_len), dim=1).to(torch.int32).contiguous()
    epc = edges_per_chunk or choose_edges_per_chunk(e / max(n_seg, 1))
    if hash_words and nb != 1:
        raise ValueError("the hashed filter needs a single column block")
    return BlockedGraph(n_rows, n_rows_pad, graph.n_sources, graph.row_offset, nb, block_size, rpt, _to_i16(row_start),
                        seg_info, col_out, e, epc, int(hash_words), w_out, graph.i_bar, graph.i_bar_agent)


STREAM_UNIT = 256                       # entries per unit of the stream layout (AT_STREAM_UNIT)
STREAM_MAX_TILE_ROWS = 2048
STREAM_MAX_BLOCK_SIZE = (1 << 21) - 1024


def choose_tile_rows(n_rows, n_blocks, avg_degree=20.0, max_padding=0.04):
    """Rows per tile of the stream layout: every (source block, tile) run is padded to whole units of 256 entries, i.e.
    by half a unit on average, so tiles are as large as the 11 row bits of an entry allow unless the shard is so small
    that one tile would leave most CTAs without work."""
    tile = STREAM_MAX_TILE_ROWS
    while tile > 32 and n_rows < tile * 8:
        tile //= 2
    return tile


def block_graph_stream(graph: ContactGraph, block_size: int, tile_rows: int = STREAM_MAX_TILE_ROWS,
                       bank_aware: bool = True) -> BlockedGraph:
    """ContactGraph -> stream layout (AT_LAYOUT_STREAM, include/at_seirm.h): for every source block ONE flat stream of
    32-bit entries, tile after tile, every (block, tile) run padded to whole 256-entry units with entries that can never
    hit.  An entry describes itself: (bitmap word of the block-local source << 16) | (row_in_tile << 5) | bit of the source,
    so the stream needs no per-row metadata and puts no limit on a row's degree.  Replaces the (2,E) COO edge_index of
    the reference (utils.py:129-148).

    The order of the entries inside a run does not matter (the kernels sum in fixed point), which `bank_aware` uses: the
    gather kernel tests the 32 entries {128 q + 4 lane + j} of a unit with ONE shared-memory load instruction, so the run
    is arranged in such groups of 32 with lane = shared-memory bank of the entry's bitmap word (word mod 32) as long as
    every bank still has entries — those loads are conflict-free — and the few entries of the over-full banks follow in
    ordinary groups."""
    dev = graph.device
    if graph.edge_weight is not None:
        raise NotImplementedError("the stream layout does not carry edge weights")
    if block_size % 1024 or not 1024 <= block_size <= STREAM_MAX_BLOCK_SIZE:
        raise ValueError("block_size must be a multiple of 1024 in 1024..2^21-1024")
    if tile_rows & (tile_rows - 1) or not 32 <= tile_rows <= STREAM_MAX_TILE_ROWS:
        raise ValueError("tile_rows must be a power of two in 32..2048")
    n_rows, e = graph.n_rows, graph.num_edges
    pad = max(tile_rows, 256)
    n_rows_pad = -(-max(n_rows, 1) // pad) * pad
    nb = max(1, -(-graph.n_sources // block_size))
    if nb > 64:
        raise ValueError("more than 64 column blocks: shard the graph over more GPUs")
    n_tiles = n_rows_pad // tile_rows
    shift = tile_rows.bit_length() - 1
    row_ptr = graph.row_ptr.long()
    deg = row_ptr[1:] - row_ptr[:-1]
    max_deg = int(deg.max()) if n_rows else 0
    # hit records of a tile: one slot per edge of its rows, so the list can never overflow
    tile_edges = torch.zeros(n_tiles * tile_rows, dtype=torch.int64, device=dev)
    tile_edges[:n_rows] = deg
    tile_edges = tile_edges.reshape(n_tiles, tile_rows).sum(1)
    tile_rec_base = (torch.cumsum(tile_edges, 0) - tile_edges).contiguous()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=dev), deg)
    del deg
    col = graph.col_idx[:e].long()
    blk = torch.div(col, block_size, rounding_mode="floor")
    n_layers = max(len(graph.i_bar_layers), 1)
    if graph.edge_layer is not None:
        blk_lay = blk * n_layers + graph.edge_layer[:e].long()       # runs per (block, layer, tile): a unit has ONE layer
    else:
        blk_lay = blk
    seg = blk_lay * n_tiles + (rows >> shift)             # run id = (block * n_layers + layer) * n_tiles + tile
    del blk_lay
    local = col - blk * block_size
    enc = ((local >> 5) << 16) | ((rows & (tile_rows - 1)) << 5) | (local & 31)
    bank = (local >> 5) & 31
    del col, blk, rows, local
    n_seg = nb * n_layers * n_tiles
    GROUP = 32
    GPU_ = STREAM_UNIT // GROUP                           # groups per unit (8)
    if bank_aware and e > 0:
        key = seg * 32 + bank                             # (run, bank) classes
        order = torch.sort(key, stable=True).indices
        key, enc = key[order], enc[order]
        del order, bank
        seg = key >> 5
        cls = torch.bincount(key, minlength=n_seg * 32)
        cls_start = torch.cumsum(cls, 0) - cls
        rank = torch.arange(e, device=dev) - cls_start[key]          # rank of the entry inside its class
        # conflict-free groups: as many as the 12th smallest class of the run has entries (below that, banks would need padding)
        thr = torch.sort(cls.reshape(n_seg, 32), dim=1).values[:, 11].contiguous()
        seg_len = cls.reshape(n_seg, 32).sum(1)
        seg_first = torch.cumsum(seg_len, 0) - seg_len
        left = rank >= thr[seg]
        cs = torch.cumsum(left.to(torch.int64), 0)
        cs_before = cs - left.to(torch.int64)
        left_idx = cs - 1 - cs_before[seg_first.clamp(max=max(e - 1, 0))][seg]      # index among the run's leftover entries
        n_left = torch.zeros(n_seg, dtype=torch.int64, device=dev)
        n_left.scatter_add_(0, seg, left.to(torch.int64))
        group = torch.where(left, thr[seg] + torch.div(left_idx, GROUP, rounding_mode="floor"), rank)
        lane = torch.where(left, left_idx % GROUP, key & 31)
        seg_groups = thr + (n_left + (GROUP - 1)) // GROUP
        seg_groups = torch.where(seg_len > 0, seg_groups, torch.zeros_like(seg_groups))
        del key, cls, cls_start, rank, left, cs, cs_before, left_idx, n_left
    else:
        del bank
        if nb > 1:
            order = torch.sort(seg, stable=True).indices  # rows are sorted already, so a run's tiles stay contiguous
            seg, enc = seg[order], enc[order]
            del order
        seg_len = torch.bincount(seg, minlength=n_seg)
        seg_first = torch.cumsum(seg_len, 0) - seg_len
        idx = torch.arange(e, device=dev) - seg_first[seg]
        group, lane = torch.div(idx, GROUP, rounding_mode="floor"), idx % GROUP
        seg_groups = (seg_len + (GROUP - 1)) // GROUP
        del idx
    seg_units = (seg_groups + (GPU_ - 1)) // GPU_
    unit_base = torch.zeros(n_seg + 1, dtype=torch.int64, device=dev)
    unit_base[1:] = torch.cumsum(seg_units, 0)
    n_units = int(unit_base[-1])
    if (n_units + 1) * STREAM_UNIT >= 2**32:
        raise ValueError("graph shard too large for 32-bit stream positions")
    # group g of a run -> unit g // 8, and inside the unit the entries {128 q + 4 lane + j}, q = (g % 8) // 4, j = g % 4:
    # the 32 entries one shared-memory load instruction of the gather kernel tests
    gi = group % GPU_
    pos = (unit_base[:-1][seg] + torch.div(group, GPU_, rounding_mode="floor")) * STREAM_UNIT + (gi // 4) * 128 + lane * 4 + (gi % 4)
    del seg, group, lane, gi
    pad_entry = (block_size >> 5) << 16                   # word block_size/32 of the staged bitmap is always zero
    pad_entry = pad_entry - 2**32 if pad_entry >= 2**31 else pad_entry
    col_out = torch.full((max(n_units, 1) * STREAM_UNIT + TAIL_PAD,), pad_entry, dtype=torch.int32, device=dev)
    enc = torch.where(enc >= 2**31, enc - 2**32, enc).to(torch.int32)     # the 32-bit pattern, as int32
    col_out[pos] = enc
    del pos, enc
    seg_ids = torch.arange(n_seg, device=dev)
    # tile in the low 24 bits, network layer in the top 8
    tile_of_seg = ((seg_ids % n_tiles) | ((torch.div(seg_ids, n_tiles, rounding_mode="floor") % n_layers) << 24)).to(torch.int32)
    unit_tile = torch.repeat_interleave(tile_of_seg, seg_units)
    if unit_tile.numel() == 0:
        unit_tile = torch.zeros(1, dtype=torch.int32, device=dev)
    if n_tiles >= 1 << 24:
        raise ValueError("too many row tiles for the 24-bit tile field")
    block_units = unit_base[torch.arange(nb + 1, device=dev) * (n_layers * n_tiles)].to(torch.int32).contiguous()
    row_start = torch.zeros(1, dtype=torch.int16, device=dev)            # unused in this layout
    g = BlockedGraph(n_rows, n_rows_pad, graph.n_sources, graph.row_offset, nb, block_size, tile_rows, row_start,
                     unit_tile.contiguous(), col_out, e, 512, 0, None, graph.i_bar, graph.i_bar_agent, 1, tile_rows)
    g.block_units = block_units
    g.tile_rec_base = tile_rec_base
    g.lam_deg_bits = max(int(max_deg).bit_length(), 1)
    g.i_bar_layers = tuple(graph.i_bar_layers)
    return g
