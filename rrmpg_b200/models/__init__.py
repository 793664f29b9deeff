"""Drop-in replacements for the models of ``rrmpg.models`` (``rrmpg/models/__init__.py:11-18``): the five
hot-path models of the north star and the snow-ice couplings."""
from .abcmodel import ABCModel
from .hbvedu import HBVEdu
from .gr4j import GR4J
from .cemaneige import Cemaneige
from .cemaneigegr4j import CemaneigeGR4J
from ._snowice import CemaneigeGR4JIce, CemaneigeHystGR4J, CemaneigeHystGR4JIce

__all__ = ["ABCModel", "HBVEdu", "GR4J", "Cemaneige", "CemaneigeGR4J", "CemaneigeHystGR4J", "CemaneigeGR4JIce",
           "CemaneigeHystGR4JIce"]
