"""fce_yolo_b200: B200-native (sm_100a) FCE-YOLOv11 detection forward path.

Importable name of the ``fce-yolo_b200`` package (a hyphen is not a valid Python identifier).
"""
__version__ = "0.1.0"

from .install import install, install_nms, uninstall, uninstall_nms  # noqa: E402,F401
from .model import YOLO  # noqa: E402,F401
