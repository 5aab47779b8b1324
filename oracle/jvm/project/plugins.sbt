addSbtPlugin("com.typesafe.sbt" % "sbt-aspectj" % "0.10.0")
