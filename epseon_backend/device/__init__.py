"""Device sub-packages (forwarders to epseon_backend_b200.device)."""
