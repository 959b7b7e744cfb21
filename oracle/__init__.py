"""CPU oracle of the SimPLe world-model hot path -- test infrastructure, never imported by simple_b200."""
