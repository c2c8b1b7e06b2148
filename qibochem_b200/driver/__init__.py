from qibochem_b200.driver.molecule import Molecule

__all__ = ["Molecule"]
