"""heat_b200 — B200-native (sm_100a) implementation of Heat's distributed k-means Lloyd path.

Public surface mirrors the reference for this one path:
``heat_b200.cluster.KMeans`` (heat/cluster/kmeans.py), ``heat_b200.spatial.cdist / rbf / manhattan``
(heat/spatial/distance.py:136-206), ``heat_b200.array`` / ``DNDarray`` (split=0 semantics), and the other consumers of
the assignment kernel: ``cluster.KMedians`` / ``KMedoids``, ``classification.KNeighborsClassifier``.
The CUDA library is loaded on first use; importing the package needs neither a GPU nor the library.
"""
from . import classification, cluster, communication, engine, spatial  # noqa: F401
from .communication import get_comm, init_from_env, use_comm  # noqa: F401
from .dndarray import DNDarray, array  # noqa: F401

float32 = __import__("torch").float32
float64 = __import__("torch").float64
int32 = __import__("torch").int32
int64 = __import__("torch").int64

__version__ = "0.1.0"
