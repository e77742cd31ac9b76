"""Connectors of the sampling-based planners (reference planning/sampling_based/__init__.py)."""
from .connectors import Connector, ConnectorParams, DiscreteConnector, ArmValidityChecker

__all__ = ["Connector", "ConnectorParams", "DiscreteConnector", "ArmValidityChecker"]
