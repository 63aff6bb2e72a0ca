"""Neutral model specifications: state_dict layouts (names / shapes) of the reference's modules and seeded synthetic
weights / audio.  No arithmetic of the hot path lives here: both the product's bench arm and the oracle import it, so the
product never has to import anything under oracle/."""
