"""Output formats the reference's tooling consumes."""
