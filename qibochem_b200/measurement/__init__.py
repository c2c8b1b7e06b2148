from qibochem_b200.measurement.result import expectation, expectation_from_samples

__all__ = ["expectation", "expectation_from_samples"]
