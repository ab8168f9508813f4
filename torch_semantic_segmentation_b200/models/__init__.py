from .fastscnn import FastSCNN, fastscnn
