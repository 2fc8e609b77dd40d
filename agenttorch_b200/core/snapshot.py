"""PackedSnapshot — an agent/environment state held as ONE pinned host buffer in compact dtypes.

`Runner.restore_state(snapshot)` re-arms an episode from a saved initial condition or checkpoint.  With a nested dict of
host tensors that is one host->device copy per tensor; a PackedSnapshot lays the tensors out back to back (16-byte aligned,
narrow dtypes: uint8 stages, int16 timers) in a single pinned buffer, so the restore is one `cudaMemcpyAsync` into a device
staging buffer followed by device-side widening copies into the live state tensors (whose storage — and therefore the
engine bindings and captured CUDA graphs — stays valid).  The reference re-runs every initialiser instead
(core/initializer.py:405-428, `reset_state`)."""
import torch


class PackedSnapshot:
    def __init__(self, tree, compact=None, pin=True):
        """tree: nested dict of tensors (device or host); compact: {leaf name: dtype} of the narrow host dtype of a leaf (the
        values must survive the round trip, which is checked)."""
        compact = compact or {}
        self.entries = []            # (path tuple, dtype, shape, byte offset, byte length)
        flat = []

        def walk(node, path):
            for k, v in node.items():
                if isinstance(v, dict):
                    walk(v, path + (k,))
                elif torch.is_tensor(v):
                    flat.append((path + (k,), v))

        walk(tree, ())
        off = 0
        staged = []
        for path, v in flat:
            dt = compact.get(path[-1], v.dtype)
            narrow = v.detach().to("cpu").to(dt).contiguous()
            if dt != v.dtype and not torch.equal(narrow.to(v.dtype), v.detach().to("cpu")):
                raise ValueError(f"{'/'.join(path)} does not survive the round trip through {dt}")
            nbytes = narrow.numel() * narrow.element_size()
            self.entries.append((path, dt, tuple(v.shape), off, nbytes))
            staged.append(narrow)
            off += (nbytes + 15) // 16 * 16
        self.nbytes = max(off, 16)
        self.host = torch.empty(self.nbytes, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            self.host = self.host.pin_memory()
        for (path, dt, shape, o, nb), narrow in zip(self.entries, staged):
            self.host[o:o + nb].copy_(narrow.view(-1).view(torch.uint8))
        self._staging = {}

    def payload_bytes(self):
        return sum(e[4] for e in self.entries)

    def restore_into(self, state, non_blocking=True):
        """One host->device copy, then one widening copy per tensor on the device."""
        first = None
        for path, *_ in self.entries:
            node = state
            for k in path:
                node = node[k]
            first = node
            break
        dev = first.device if first is not None else torch.device("cpu")
        stage = self._staging.get(dev)
        if stage is None:
            stage = self._staging[dev] = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        stage.copy_(self.host, non_blocking=non_blocking)
        for path, dt, shape, off, nb in self.entries:
            node = state
            for k in path[:-1]:
                node = node[k]
            cur = node[path[-1]]
            src = stage[off:off + nb].view(dt).reshape(shape)
            if cur.requires_grad or cur.grad_fn is not None or tuple(cur.shape) != tuple(shape):
                node[path[-1]] = src.to(cur.dtype).reshape(cur.shape).clone()
            else:
                cur.copy_(src)
