"""Test-infrastructure stand-in for the subset of torch_geometric that the
reference's SEIRM path touches (torch_geometric is not installed in this image
and there is no network).  Only `MessagePassing.propagate` with
flow="source_to_target", aggr="add" and the `Data` attribute bag are provided;
semantics follow PyG's documentation: `foo_i = foo[edge_index[1]]`,
`foo_j = foo[edge_index[0]]`, messages are summed onto `edge_index[1]`.
NOT product code: used only by oracle/ref_harness to run /root/reference."""
