from .cross_entropy import CrossEntropyLoss, cross_entropy
from .ohem_loss import OHEMLoss, ohem_loss
