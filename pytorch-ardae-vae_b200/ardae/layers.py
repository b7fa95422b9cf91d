"""Parameter containers with the reference's module / state_dict structure.

`MLP` mirrors models/layers.py:477-499 (`layers.{i}` + `fc`, nn.Linear default init) and
`ContextConcatMLP` mirrors models/layers.py:681-705.  They hold parameters only: the arithmetic of
the hot path runs in the fused CUDA plans (libardae), never through per-layer torch ops.
"""
import torch.nn as nn


class MLP(nn.Module):
    def __init__(self, input_dim=2, hidden_dim=8, output_dim=2, nonlinearity='relu', num_hidden_layers=1,
                 use_nonlinearity_output=False):
        super().__init__()
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.nonlinearity, self.num_hidden_layers = nonlinearity, num_hidden_layers
        self.use_nonlinearity_output = use_nonlinearity_output
        self.layers = nn.ModuleList([nn.Linear(input_dim if i == 0 else hidden_dim, hidden_dim)
                                     for i in range(num_hidden_layers)])
        self.fc = nn.Linear(input_dim if num_hidden_layers == 0 else hidden_dim, output_dim)

    def linears(self):
        return list(self.layers) + [self.fc]

    def forward(self, *a, **k):
        raise RuntimeError('ardae.layers.MLP is a parameter container; the fused CUDA plan of the owning '
                           'module evaluates it')


class ContextConcatMLP(nn.Module):
    def __init__(self, input_dim=2, context_dim=2, hidden_dim=8, output_dim=2, nonlinearity='relu',
                 num_hidden_layers=1, use_nonlinearity_output=False):
        super().__init__()
        self.input_dim, self.context_dim = input_dim, context_dim
        self.hidden_dim, self.output_dim = hidden_dim, output_dim
        self.nonlinearity, self.num_hidden_layers = nonlinearity, num_hidden_layers
        self.use_nonlinearity_output = use_nonlinearity_output
        self.layers = nn.ModuleList([nn.Linear((input_dim if i == 0 else hidden_dim) + context_dim, hidden_dim)
                                     for i in range(num_hidden_layers)])
        self.fc = nn.Linear((input_dim if num_hidden_layers == 0 else hidden_dim) + context_dim, output_dim)

    def linears(self):
        return list(self.layers) + [self.fc]

    def forward(self, *a, **k):
        raise RuntimeError('ardae.layers.ContextConcatMLP is a parameter container')
