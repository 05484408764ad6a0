"""Core data model, permutation-spec compiler and LAP solver entry point."""
