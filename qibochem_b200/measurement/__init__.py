from qibochem_b200.measurement.result import expectation, expectation_from_samples, sample_statistics, v_expectation

__all__ = ["expectation", "expectation_from_samples", "sample_statistics", "v_expectation"]
