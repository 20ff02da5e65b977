//! serde codecs of the wire format (reference src/serialize.rs:1-78): a BigInt is a DECIMAL string, a Vec<BigInt> a
//! sequence of decimal strings.  Fields without a `with =` attribute in the reference keep curv's own BigInt serde.
pub mod bigint {
    use curv::arithmetic::traits::*;
    use curv::BigInt;
    use serde::{de, ser, Deserialize};
    pub fn serialize<S: ser::Serializer>(x: &BigInt, s: S) -> Result<S::Ok, S::Error> {
        s.serialize_str(&x.to_str_radix(10))
    }
    pub fn deserialize<'de, D: de::Deserializer<'de>>(d: D) -> Result<BigInt, D::Error> {
        let s = String::deserialize(d)?;
        BigInt::from_str_radix(&s, 10).map_err(de::Error::custom)
    }
}
pub mod vecbigint {
    use curv::arithmetic::traits::*;
    use curv::BigInt;
    use serde::ser::SerializeSeq;
    use serde::{de, ser, Deserialize};
    pub fn serialize<S: ser::Serializer>(x: &[BigInt], s: S) -> Result<S::Ok, S::Error> {
        let mut seq = s.serialize_seq(Some(x.len()))?;
        for e in x {
            seq.serialize_element(&e.to_str_radix(10))?;
        }
        seq.end()
    }
    pub fn deserialize<'de, D: de::Deserializer<'de>>(d: D) -> Result<Vec<BigInt>, D::Error> {
        let v = Vec::<String>::deserialize(d)?;
        // a malformed digit string panics in the reference (`unwrap()`, serialize.rs:69): kept
        Ok(v.iter().map(|s| BigInt::from_str_radix(s, 10).unwrap()).collect())
    }
}
