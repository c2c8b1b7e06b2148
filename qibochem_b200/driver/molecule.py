"""
``Molecule``: holder of the molecular integrals and producer of the Hamiltonian term list.

Reference: ``src/qibochem/driver/molecule.py`` -- ``Molecule`` :22-91, ``hamiltonian`` :398-466.  The classical
chemistry behind the integrals (PySCF RHF, ao2mo, MP2 natural orbitals, ``run_pyscf`` :164-221) is outside the B200
hot path: integrals are supplied by the caller (``oei``, ``tei``, ``e_nuc``, ``eps``) or, when PySCF happens to be
installed, by ``run_pyscf``.  ``hamiltonian()`` keeps the reference's signature and returns a mask-array
``SymbolicHamiltonian`` (or the intermediate fermionic / qubit operators).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from qibochem_b200.driver.hamiltonian import fermionic_terms, qubit_hamiltonian
from qibochem_b200.hamiltonian import SymbolicHamiltonian


class Molecule:
    def __init__(self, geometry=None, charge: int = 0, multiplicity: int = 1, basis: str | None = None,
                 xyz_file: str | None = None, active=None):  # fmt: skip
        if geometry is None and xyz_file is not None:
            self._process_xyz_file(xyz_file, charge, multiplicity)
        else:
            self.geometry = geometry
            self.charge = charge
            self.multiplicity = multiplicity
        self.basis = "sto-3g" if basis is None else basis
        self.nelec = None  # number of electrons
        self.norb = None  # number of spatial orbitals
        self.nso = None  # number of spin-orbitals
        self.e_hf = None
        self.oei = None  # one-electron integrals, MO basis
        self.tei = None  # two-electron integrals, MO basis, second-quantisation index order
        self.ca = None
        self.pa = None
        self.da = None
        self.eps = None  # orbital energies
        self.e_nuc = None
        self.active = active
        self.frozen = None
        self.inactive_energy = None
        self.embed_oei = None
        self.embed_tei = None
        self.n_active_e = None
        self.n_active_orbs = None

    def _process_xyz_file(self, xyz_file, charge, multiplicity):
        """xyz file -> geometry; an optional "charge multiplicity" pair on the comment line overrides the arguments
        (molecule.py:98-123)."""
        assert Path(f"{xyz_file}").exists(), f"{xyz_file} not found!"
        with open(xyz_file, encoding="utf-8") as handle:
            lines = handle.readlines()
        n_atoms = int(lines[0])
        if lines[1].split():
            charge, multiplicity = (int(v) for v in lines[1].split())
        self.geometry = [(ln.split()[0], tuple(float(v) for v in ln.split()[1:4])) for ln in lines[2 : n_atoms + 2]]
        self.charge, self.multiplicity = charge, multiplicity

    @classmethod
    def from_integrals(cls, oei, tei, nelec: int, e_nuc: float = 0.0, eps=None) -> "Molecule":
        """Molecule defined directly by MO integrals (what ``run_pyscf`` would have produced)."""
        mol = cls()
        mol.oei = np.asarray(oei, dtype=float)
        mol.tei = np.asarray(tei, dtype=float)
        mol.norb = mol.oei.shape[0]
        mol.nso = 2 * mol.norb
        mol.nelec = int(nelec)
        mol.e_nuc = float(e_nuc)
        mol.eps = np.asarray(eps, dtype=float) if eps is not None else np.diag(mol.oei).copy()
        return mol

    def run_pyscf(self, max_scf_cycles: int = 50, do_mp2: bool = False):
        """Classical RHF with PySCF (molecule.py:164-221).  Not part of the B200 path; needs PySCF installed."""
        try:
            import pyscf
        except ImportError as exc:  # pragma: no cover
            raise ImportError("PySCF is not installed: supply integrals with Molecule.from_integrals(...)") from exc
        mol = pyscf.gto.M(atom=self.geometry, basis=self.basis, charge=self.charge, spin=self.multiplicity - 1)
        hf = pyscf.scf.RHF(mol)
        hf.max_cycle = max_scf_cycles
        hf.run()
        self.e_hf = hf.e_tot
        self.nelec = sum(mol.nelec)
        self.ca = np.asarray(hf.mo_coeff)
        self.eps = np.asarray(hf.mo_energy)
        self.norb = self.ca.shape[1]
        self.nso = 2 * self.norb
        self.oei = self.ca.T @ hf.get_hcore() @ self.ca
        eri = pyscf.ao2mo.kernel(mol, self.ca)
        eri4 = pyscf.ao2mo.restore("s1", eri, self.norb)
        self.tei = np.einsum("pqrs->prsq", eri4)
        self.e_nuc = mol.energy_nuc()

    @staticmethod
    def _filter_array(array, threshold):
        out = np.copy(array)
        out[np.abs(array) < threshold] = 0.0
        return out

    def hamiltonian(self, ham_type=None, oei=None, tei=None, constant=None, ferm_qubit_map=None, threshold=1e-12):
        """Molecular Hamiltonian from the one-/two-electron integrals.

        Args:
            ham_type: ``"f"/"ferm"`` -> fermionic terms dict, ``"q"/"qubit"`` -> ``PauliSum``,
                ``"s"/"sym"`` (default) -> ``SymbolicHamiltonian``
            oei, tei: integrals (default: the embedded ones if HF embedding was applied, else the full ones)
            constant: added to the electronic energy (default: inactive Fock energy if embedded, else 0)
            ferm_qubit_map: ``"jw"`` (default) or ``"bk"``
            threshold: integrals below it are dropped
        """
        if ham_type is None:
            ham_type = "sym"
        if oei is None:
            oei = self.oei if self.embed_oei is None else self.embed_oei
        if tei is None:
            tei = self.tei if self.embed_tei is None else self.embed_tei
        if constant is None:
            constant = 0.0 if self.inactive_energy is None else self.inactive_energy
        if ferm_qubit_map is None:
            ferm_qubit_map = "jw"
        constant = constant + self.e_nuc
        oei = self._filter_array(oei, threshold)
        tei = self._filter_array(tei, threshold)
        if ham_type in ("f", "ferm"):
            return fermionic_terms(oei, tei, constant)
        qubit_op = qubit_hamiltonian(oei, tei, constant, ferm_qubit_map)
        if ham_type in ("q", "qubit"):
            return qubit_op
        if ham_type in ("s", "sym"):
            return SymbolicHamiltonian(qubit_op, nqubits=2 * np.asarray(oei).shape[0])
        raise NameError(f"Unknown {ham_type}!")
