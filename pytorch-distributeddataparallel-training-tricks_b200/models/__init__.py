"""Model families: Toy_Net (the reference's CNN) and ResNet-50 (BASELINE.json config 4)."""
from .toy_net import Toy_Net, TOY_NET_CONV_PLAN, toy_net_flops_per_image
from .resnet import ResNet50, resnet50

__all__ = ["Toy_Net", "TOY_NET_CONV_PLAN", "toy_net_flops_per_image", "ResNet50", "resnet50"]
