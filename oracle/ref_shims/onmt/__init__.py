"""Test-infrastructure shim: a restatement of the few OpenNMT-py==2.2.0 classes the
reference decoder imports (MolNexTR/models/decoder.py:9-13, models/embedding.py:8).
OpenNMT-py is not vendored in /root/reference and not installed here; these files
restate its published algorithm so the reference's OWN modules can be imported in this
container to pin the oracle (see oracle/README.md).  Never imported by the product."""
