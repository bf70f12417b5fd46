from .FusedLossComputer01 import FusedLossComputer, ray_losses, stream_plan  # noqa: F401
