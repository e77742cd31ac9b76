"""Connectors of the sampling-based planners (reference planning/sampling_based/__init__.py)."""
from .connectors import (Connector, ConnectorParams, DiscreteConnector, ContinuousConnector, ArmValidityChecker,
                         ArmClearanceChecker)
from .roadmap import BatchedPRM, Roadmap, knn_earlier

__all__ = ["Connector", "ConnectorParams", "DiscreteConnector", "ContinuousConnector", "ArmValidityChecker", "ArmClearanceChecker", "BatchedPRM", "Roadmap", "knn_earlier"]
