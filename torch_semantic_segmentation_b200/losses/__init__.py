from .cross_entropy import CrossEntropyLoss, cross_entropy
