"""Codec, checkpoint and metrics utilities."""
