"""Katib ``api.v1.beta1`` Suggestion gRPC surface (GetSuggestions / ValidateAlgorithmSettings) backed by libkbo."""
