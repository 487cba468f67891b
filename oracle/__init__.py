"""CPU oracle for the Zephyr scoring path -- test infrastructure, never imported by the product."""
