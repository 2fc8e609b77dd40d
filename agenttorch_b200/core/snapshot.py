"""PackedSnapshot — an agent/environment state held as ONE pinned host buffer in compact dtypes.

`Runner.restore_state(snapshot)` re-arms an episode from a saved initial condition or checkpoint.  With a nested dict of
host tensors that is one host->device copy per tensor; a PackedSnapshot lays the tensors out back to back (16-byte aligned,
narrow dtypes: uint8 stages, int16 timers) in a single pinned buffer, so the restore is one `cudaMemcpyAsync` into a device
staging buffer followed by device-side widening copies into the live state tensors (whose storage — and therefore the
engine bindings and captured CUDA graphs — stays valid).  `prefetch(device)` starts that copy early, on a copy stream and into
the second of two staging buffers, so that the next restore finds its bytes on the device (an input pipeline: episode k+1's
initial condition crosses PCIe while episode k computes).  The reference re-runs every initialiser instead
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

    def _device_slot(self, dev):
        slot = self._staging.get(dev)
        if slot is None:
            slot = self._staging[dev] = {"bufs": [None, None], "consumed": [None, None], "ready": None, "next": 0, "stream": None}
        return slot

    def _stage_buffer(self, slot, i, dev):
        if slot["bufs"][i] is None:
            slot["bufs"][i] = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        return slot["bufs"][i]

    def prefetch(self, device):
        """Start the host->device copy for the NEXT restore_into on `device` (copy stream, alternate staging buffer).  No-op on
        the CPU; calling it twice without a restore in between keeps the first copy."""
        dev = torch.device(device)
        if dev.type != "cuda":
            return
        slot = self._device_slot(dev)
        if slot["ready"] is not None:
            return
        i = slot["next"]
        buf = self._stage_buffer(slot, i, dev)
        if slot["stream"] is None:
            slot["stream"] = torch.cuda.Stream(device=dev)
        cs = slot["stream"]
        if slot["consumed"][i] is not None:
            cs.wait_event(slot["consumed"][i])          # the widening copies that last read this staging buffer
        with torch.cuda.stream(cs):
            buf.copy_(self.host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        slot["ready"] = (i, ev)

    def restore_into(self, state, non_blocking=True):
        """One host->device copy (already under way if prefetch() was called), then one widening copy per tensor on the device."""
        first = None
        for path, *_ in self.entries:
            node = state
            for k in path:
                node = node[k]
            first = node
            break
        dev = first.device if first is not None else torch.device("cpu")
        slot = self._device_slot(dev)
        if slot["ready"] is not None:
            i, ev = slot["ready"]
            slot["ready"] = None
            torch.cuda.current_stream(dev).wait_event(ev)
            stage = slot["bufs"][i]
        else:
            i = slot["next"]
            stage = self._stage_buffer(slot, i, dev)
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
        if dev.type == "cuda":
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
            slot["consumed"][i] = done
        slot["next"] = 1 - i
