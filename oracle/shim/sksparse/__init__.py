"""Stand-in for scikit-sparse (CHOLMOD), reference gpr.py:10,98-108,342-352.  TEST INFRASTRUCTURE ONLY."""
