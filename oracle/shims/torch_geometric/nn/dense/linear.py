"""PyG 2.1.0 nn.dense.linear.Linear: F.linear with weight [out,in]; default init is
kaiming-uniform(a=sqrt(5)) on the weight and U(+-1/sqrt(in)) on the bias -- i.e. the
torch.nn.Linear default, so state_dict keys/shapes are interchangeable."""
import torch


class Linear(torch.nn.Linear):
    def __init__(self, in_channels, out_channels, bias=True,
                 weight_initializer=None, bias_initializer=None):
        super().__init__(in_channels, out_channels, bias=bias)
        self.in_channels = in_channels
        self.out_channels = out_channels
