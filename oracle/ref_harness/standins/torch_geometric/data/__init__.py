class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, **kwargs):
        self.x, self.edge_index, self.edge_attr = x, edge_index, edge_attr
        for k, v in kwargs.items():
            setattr(self, k, v)
