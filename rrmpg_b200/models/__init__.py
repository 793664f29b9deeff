"""Drop-in replacements for the five hot-path models of ``rrmpg.models`` (``rrmpg/models/__init__.py:11-18``).

The Hysteresis / Ice variants of the reference are outside this build's scope.
"""
from .abcmodel import ABCModel
from .hbvedu import HBVEdu
from .gr4j import GR4J
from .cemaneige import Cemaneige
from .cemaneigegr4j import CemaneigeGR4J

__all__ = ["ABCModel", "HBVEdu", "GR4J", "Cemaneige", "CemaneigeGR4J"]
