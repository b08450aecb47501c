"""Command-line debugging tools."""
