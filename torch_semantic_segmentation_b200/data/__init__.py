from .cityscapes import CLASSES, TRAIN_MAPPING, DeviceTransform, eval_transform, train_transform

__all__ = ['CLASSES', 'TRAIN_MAPPING', 'DeviceTransform', 'train_transform', 'eval_transform']
