"""Test infrastructure only (CPU oracle). Never imported by refid_b200/."""
