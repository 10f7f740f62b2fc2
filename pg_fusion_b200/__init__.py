"""pg_fusion_b200: B200-native (sm_100a) worker-side columnar hot path of pg_fusion.

The package is a thin host layer over libpgf_b200.so (hand-written CUDA behind a C ABI,
include/pgf_b200.h).  It mirrors the reference's interfaces for this path:
`arrow_layout` (page format), `runtime_filter` (Bloom filter + lifecycle) and the fused
operator pipelines the worker plans over a scan.  No CPU fallback exists.
"""
from . import arrow_layout, errors
from .arrow_layout import ColumnSpec, LayoutPlan, TypeTag
from .errors import PgfError
from .worker import (AggFunc, BloomParams, Cmp, Context, Factor, GenTable, GUC_DEFAULT_BLOOM, PipelineBuilder,
                     PipelineResult, ProbeDecision, RuntimeFilter, RuntimeFilterState, Scan, device_count)

__all__ = [
    "arrow_layout", "errors", "ColumnSpec", "LayoutPlan", "TypeTag", "PgfError", "AggFunc", "BloomParams", "Cmp",
    "Context", "Factor", "GenTable", "GUC_DEFAULT_BLOOM", "PipelineBuilder", "PipelineResult", "ProbeDecision",
    "RuntimeFilter", "RuntimeFilterState", "Scan", "device_count",
]
