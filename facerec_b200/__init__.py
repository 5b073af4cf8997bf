"""facerec_b200 — B200-native per-frame face path for the MoMaF facerec pipeline.

Host side (Python) mirrors the reference's operator interface for this path — `detect()`
(facerec/detector.py:66-79), `embeddings()` / `cropBox` (facerec/extract.py:161-167), the linkage
distance input (facerec/cluster.py:41) and the kNN classifier (facerec/classify_knn.py:212,227) —
on top of the C ABI of libfacerec_b200.so (include/facerec_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
