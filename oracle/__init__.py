"""CPU oracle of the AquaLoRA hot paths -- TEST INFRASTRUCTURE, never imported by aqualora_b200/.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
"""
