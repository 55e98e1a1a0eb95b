"""radardistill_b200 -- B200 (sm_100a) dynamic pillar encoder behind RadarDistill's VFE plugin API.

Public surface (mirrors ``pcdet/models/backbones_3d/vfe``):
    DynamicPillarVFE, DynamicPillarVFESimple2D, Radar_DynamicPillarVFESimple2D,
    Radar_DynamicPillarVFESimple2D_Test, PFNLayerV2, VFETemplate, REGISTRY, register
The computation lives in ``librdp.so`` (C ABI: ``include/rdp.h``), built by ``radardistill_b200.build``.
"""
from .vfe import (REGISTRY, DynamicPillarVFE, DynamicPillarVFESimple2D, PFNLayerV2, Radar_DynamicPillarVFESimple2D,
                  Radar_DynamicPillarVFESimple2D_Test, VFETemplate, forward_pair, register)
from . import ops, synth  # noqa: F401

__version__ = "0.1.0"
