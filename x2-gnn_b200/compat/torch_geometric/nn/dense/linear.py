"""torch_geometric.nn.dense.linear.Linear: torch.nn.Linear with PyG's constructor keywords; the default
initialisation (kaiming-uniform a=sqrt(5), bias U(+-1/sqrt(in))) and the state_dict keys are the same."""
import torch


class Linear(torch.nn.Linear):
    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__(in_channels, out_channels, bias=bias)
        if weight_initializer == "glorot":
            torch.nn.init.xavier_uniform_(self.weight)
        if bias and bias_initializer == "zeros":
            torch.nn.init.zeros_(self.bias)
