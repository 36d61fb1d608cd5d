"""Mirror of the reference package `core` (lib/core): only the test-time detection post-processing lives here."""
