from .FusedLossComputer01 import FusedLossComputer, ray_losses, reprojection_losses, stream_plan, view_tables  # noqa: F401
