"""hcir_b200 -- B200-native exact cosine top-k / kNN-vote path of
atunnd/Hair-centric-Image-Retrieval (HSimCLR), behind the reference's Python call surface.

Import name: ``hcir_b200`` (the directory is ``hair-centric-image-retrieval_b200``; the
repo-root ``hcir_b200.py`` shim registers it under the importable name)."""
from .engine import (GalleryBank, FeatureBankBuilder, PendingStep, knn_topk, knn_predict,  # noqa: F401
                     l2_normalize)
from .classifier import KNeighborsClassifierB200  # noqa: F401
from .retrieval import (retrieve_similar_images, HairRetrievalB200, FlatIndex,  # noqa: F401
                        compute_similarity_topk, clear_bank_cache)
from .sharded import (ShardPlan, ShardedGallery, QueryShardedGallery, choose_sharding,  # noqa: F401
                      exchange_candidates)
from .pipeline import HostPipeline, HostPending  # noqa: F401
from . import synth, formats, metrics  # noqa: F401

__all__ = [
    "GalleryBank", "FeatureBankBuilder", "PendingStep", "knn_topk", "knn_predict", "l2_normalize", "KNeighborsClassifierB200",
    "retrieve_similar_images", "HairRetrievalB200", "FlatIndex", "compute_similarity_topk",
    "clear_bank_cache", "ShardPlan", "ShardedGallery", "QueryShardedGallery", "choose_sharding", "exchange_candidates", "HostPipeline", "HostPending", "synth", "formats", "metrics",
]
