import inspect

import torch
import torch.nn as nn


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        num_nodes = None
        call = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                num_nodes = base.size(0)
                call[name] = base.index_select(0, dst if name.endswith("_i") else src)
            else:
                call[name] = kwargs[name]
        msg = self.message(**call)
        out = torch.zeros((num_nodes,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, dst, msg)

    def message(self, x_j):
        return x_j
