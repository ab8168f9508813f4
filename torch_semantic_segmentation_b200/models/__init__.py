from .fastscnn import FastSCNN, fastscnn
from .contextnet import ContextNet, contextnet12, contextnet14, contextnet18
