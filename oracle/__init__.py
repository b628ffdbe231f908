"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatements of the reference hot path).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package eventclip_b200/ must never import this package.
"""
