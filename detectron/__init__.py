"""Namespace for the B200-native re-implementation of Detectron.pytorch's RoI/NMS hot path.

The product package is :mod:`detectron.pytorch_b200`.
"""
